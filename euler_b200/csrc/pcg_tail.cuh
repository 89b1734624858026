// euler_b200/csrc/pcg_tail.cuh — the second half of a red-black PCG iteration as ONE kernel:
//
//   r' = r - alpha A s ; p += alpha s ; ||r'||inf           fmadd x2 + inf_norm   main.c:753-756
//   q  = L^-1 r'                                            red-black forward solve
//   z  = L^-T q ; z.r'                                      red-black backward solve + dot, main.c:760-762
//   beta = z.r' / sigma ; sigma = z.r' ; stop test          main.c:756-765
//
// (included by pcg_kernels.cu, inside its anonymous namespace).  The three steps used to be three
// launches (k_axpy, k_rb_forward_pipe, k_rb_backward_pipe) that moved 40 + 25 + 33 = 98 B/cell:
// r' and q were written to HBM only to be read back by the next kernel.  Here a block walks its
// row range ONCE and the intermediates never leave the SM:
//
//   * inputs (r, A s, the preconditioner diagonal, the fluid mask) arrive row by row through the
//     TMA bulk-copy ring of pcg_pipe.cuh, each row segment with 2 halo columns per side;
//   * the three steps run as a software pipeline skewed by one row each: in step yy the block
//     forms r'(yy) [stage 1], q(yy-1) [stage 2] and z(yy-2) [stage 3].  What a cell needs from its
//     four neighbours — w = pc (r' pc) of a red neighbour for the forward solve, zb = q pc of a
//     black neighbour for the backward solve — goes through ONE rolling shared-memory row buffer
//     (red cells hold w, black cells zb: a cell is never asked for the other one); a thread's own
//     r', pc and q of the two previous rows stay in registers;
//   * a row range needs r' on its +-2 rows / columns and q on +-1 (red z <- black q <- red r'):
//     the 4 + 2 halo rows are recomputed (their inputs are L2 hits: the neighbouring block loads
//     the same rows), the 4 halo columns are the job of an extra warp that also runs the ring;
//   * r is read from one plane and r' written to its twin (neighbouring blocks re-derive r' on
//     their halo rows from the OLD r), p is updated in place on the rows a block owns.
//
// Algorithmic bytes: R r, A s, pc 24 + fluid 1 + W r' 8 + W z 8 = 41 B/cell on odd iterations,
// + R s', s, p 24 + W p 8 = 73 on even ones (the deferred p update, see k_axpy): 57 B/cell mean
// instead of 98; a whole iteration 34 + 57 = 91 B/cell instead of 132.
//
// Every cell value is produced by exactly the operations, in exactly the order, of the three
// kernels it replaces (RbForwardPipe / RbBackwardPipe in pcg_ops.cuh, k_axpy): r', p, z are
// bit-identical to theirs, checked through the C-ABI stage EULER_S_FUSED_TAIL against the CPU
// mirror (tests/test_gpu_stages.py).
#pragma once

namespace tail {

constexpr int HALO = 2;                 // halo columns of the value planes (== pipe::Elem<double>::HX)
constexpr int ABW = TW + 2 * HALO;      // doubles per row of the rolling neighbour buffer
constexpr int MW = TW + 2 * pipe::HX1;  // bytes per row of the rolling mask buffer
constexpr int NSLOT = 5;                // steps j-3 .. j are live; a fast warp may already write step j+1

template <int NS>
constexpr int smem_bytes() {
  return NS * pipe::Layout<3, 1>::stage_bytes + 2 * NS * 8 + NSLOT * ABW * 8 + NSLOT * MW;
}

}  // namespace tail

// mode: 0 = r only (odd iterations), 1 = r and both pending p updates (even iterations),
//       2 = r and this iteration's p update (every iteration; not used by the fused solve)
template <int NS, int C>
__global__ void __launch_bounds__(TW / C + 32, 3) k_fused_tail(
    Grid g, TileList active, const double* __restrict__ r, const double* __restrict__ as,
    const double* __restrict__ precon, const uint8_t* __restrict__ fluid, const double* __restrict__ s,
    const double* __restrict__ s_prev, double* __restrict__ p, double* __restrict__ r_new,
    double* __restrict__ z, double* partials, DevScalars* sc, double tol, int mode, int exact,
    int acc0, int acc1, const __grid_constant__ DistArgs dist) {
  using L = pipe::Layout<3, 1>;
  constexpr int NMAIN = TW / C;                  // main threads: C cells each
  constexpr int ROWT = L::ROWT;
  if (sc->done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* stages = smem_raw;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + NS * L::stage_bytes);
  uint64_t* empty = full + NS;
  double* ab = reinterpret_cast<double*>(empty + NS);                       // [NSLOT][ABW], index col + HALO
  uint8_t* mk = reinterpret_cast<uint8_t*>(ab + tail::NSLOT * tail::ABW);   // [NSLOT][MW], index col + HX1

  const int tid = threadIdx.x, lane = tid & 31;
  const bool is_main = tid < NMAIN;
  const bool producer = tid == NMAIN;            // lane 0 of the extra warp runs the ring
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { pipe::mbar_init(full + i, 1); pipe::mbar_init(empty + i, NMAIN / 32 + 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const double alpha = sc->alpha, alpha_prev = sc->alpha_prev;
  const double neg_alpha = -alpha;
  const int th = g.th;
  const pipe::Tiles T = pipe::tiles_of(g, th);
  pipe::JobIter cons, prod;
  cons.rad = 2;
  cons.start(g, T, th, active.list, (int)*active.count);
  prod = cons;
  int issued = 0;

  auto issue = [&]() {                           // producer lane only
    const int st = issued % NS, use = issued / NS;
    pipe::mbar_wait(empty + st, (use & 1) ^ 1);
    unsigned char* dst = stages + st * L::stage_bytes;
    const uint32_t b8 = (uint32_t)(prod.p.w + 2 * tail::HALO) * 8u, b1 = (uint32_t)(prod.p.w + 2 * pipe::HX1);
    pipe::mbar_expect_tx(full + st, 3 * b8 + b1);
    const long row = (long)prod.yy * g.pitch + prod.p.x0;
    pipe::bulk_g2s(dst, r + row - tail::HALO, b8, full + st);
    pipe::bulk_g2s(dst + ROWT, as + row - tail::HALO, b8, full + st);
    pipe::bulk_g2s(dst + 2 * ROWT, precon + row - tail::HALO, b8, full + st);
    pipe::bulk_g2s(dst + 3 * ROWT, fluid + row - pipe::HX1, b1, full + st);
    ++issued;
    prod.next(g, T, th, active.list);
  };

  // columns this thread works on, relative to the piece: main threads C cells from tid*C; the
  // extra warp's lanes 0..3 one halo column each (-2, -1, w, w+1: set per piece)
  const int t0 = tid * C;
  double r1[C], pc1[C], r2[C], pc2[C], q1[C], q2[C];      // rows yy-1 / yy-2 of this thread's columns
  double acc = 0.0, mx = 0.0;
  bool peer_stored = false;
  double* __restrict__ z_dn = dist.z_dn;
  double* __restrict__ z_up = dist.z_up;
#pragma unroll
  for (int k = 0; k < C; ++k) r1[k] = pc1[k] = r2[k] = pc2[k] = q1[k] = q2[k] = 0.0;

  for (int j = 0; cons.valid; ++j) {
    if (producer)
      while (prod.valid && issued <= j + NS - 1) issue();
    const pipe::Piece pz = cons.p;
    const int yy = cons.yy;                      // row whose inputs arrive in this step
    // slots of the rolling buffers go by STEP, not by row (rows jump at a piece boundary).  No
    // barrier separates stage 3 of one step from stage 1 of the next: with 5 slots the one a fast
    // warp overwrites (step j-5's) is one nobody still reads (stage 3 reads steps j-3 .. j-1)
    const int s0 = j % tail::NSLOT, sm1 = (j + 4) % tail::NSLOT, sm2 = (j + 3) % tail::NSLOT,
              sm3 = (j + 2) % tail::NSLOT;
    double* ab0 = ab + s0 * tail::ABW + tail::HALO;     // rows yy, yy-1, yy-2, yy-3 of the neighbour buffer
    double* ab1 = ab + sm1 * tail::ABW + tail::HALO;
    double* ab2 = ab + sm2 * tail::ABW + tail::HALO;
    double* ab3 = ab + sm3 * tail::ABW + tail::HALO;
    uint8_t* mk0 = mk + s0 * tail::MW + pipe::HX1;
    const uint8_t* mk1 = mk + sm1 * tail::MW + pipe::HX1;
    const uint8_t* mk2 = mk + sm2 * tail::MW + pipe::HX1;
    const uint8_t* mk3 = mk + sm3 * tail::MW + pipe::HX1;
    // first column and number of columns of this thread in this piece
    int col, ncol;
    if (is_main) { col = t0; ncol = t0 < pz.w ? C : 0; }
    else { col = lane < 2 ? lane - 2 : pz.w + lane - 2; ncol = lane < 4 ? 1 : 0; }
    const bool own_cols = is_main && ncol > 0;
    const int gx = pz.x0 + col;                  // global column (the red/black parity needs it)
    const int gyy = yy + g.yoff;

    // the element-wise operands of the p update, issued before the wait on the ring
    double sv[C], spv[C], pv[C];
    const bool own_row = yy >= pz.y0 && yy < pz.y1;
    const size_t c0 = gidx(g, pz.x0 + t0, yy);
    const bool do_p = mode != 0 && own_row && own_cols;
    if (do_p) {
      const DV<C> a = ldg_v<C>(s + c0), b = ldg_v<C>(p + c0);
#pragma unroll
      for (int k = 0; k < C; ++k) { sv[k] = a.v[k]; pv[k] = b.v[k]; }
      if (mode == 1) {
        const DV<C> d = ldg_v<C>(s_prev + c0);
#pragma unroll
        for (int k = 0; k < C; ++k) spv[k] = d.v[k];
      }
    }

    // ---- stage 1: r'(yy), w(yy) ---------------------------------------------------------
    pipe::mbar_wait(full + (j % NS), (j / NS) & 1);
    double r0[C], pc0[C];
    {
      const pipe::RowView<3, 1> in = L::view(stages + (j % NS) * L::stage_bytes);
#pragma unroll
      for (int k = 0; k < C; ++k) {
        r0[k] = pc0[k] = 0.0;
        if (k >= ncol) continue;
        const uint8_t m = in.b[0][col + k];
        mk0[col + k] = m;
        const double rr = in.d[0][col + k], aa = in.d[1][col + k], pp = in.d[2][col + k];
        const double rn = m ? rr + aa * neg_alpha : rr;            // fmadd(z, -alpha, r), main.c:754
        r0[k] = rn; pc0[k] = pp;
        if (((gx + k + gyy) & 1) == 0) ab0[col + k] = pp * (rn * pp);   // red: what a black neighbour adds
        if (own_row && own_cols && m) {
          if (mode == 1) pv[k] = pv[k] + spv[k] * alpha_prev;      // the previous iteration's main.c:753
          if (mode) pv[k] = pv[k] + sv[k] * alpha;                 // fmadd(s, alpha, p), main.c:753
          const double a = fabs(rn);
          if (a > mx && yy >= acc0 && yy < acc1) mx = a;           // NaN-dropping max, main.c:659-662
        }
      }
      // the mask bytes left and right of the value halo are never read; the ring stage is free
      __syncwarp();
      if (lane == 0) pipe::mbar_arrive(empty + (j % NS));
      if (own_row && own_cols) {
        DV<C> o;
#pragma unroll
        for (int k = 0; k < C; ++k) o.v[k] = r0[k];
        stv<C>(r_new + c0, o);
        if (do_p) {
#pragma unroll
          for (int k = 0; k < C; ++k) o.v[k] = pv[k];
          stv<C>(p + c0, o);
        }
      }
    }
    __syncthreads();

    // ---- stage 2: q(yy-1), zb(yy-1) -----------------------------------------------------
    const int y1r = yy - 1;
    double q0[C];
#pragma unroll
    for (int k = 0; k < C; ++k) q0[k] = 0.0;
    if (y1r >= pz.y0 - 1 && y1r <= pz.y1 && yy >= pz.y0 - 1) {
      // (the extra warp: only the inner halo columns -1 and w)
      const bool active2 = is_main ? ncol > 0 : (lane == 1 || lane == 2);
      if (active2) {
#pragma unroll
        for (int k = 0; k < C; ++k) {
          if (k >= ncol) continue;
          const int cc = col + k;
          if (!mk1[cc]) continue;
          double t = r1[k];
          const bool black = ((gx + k + gyy - 1) & 1) != 0;
          if (black) {                                             // + red neighbours: l, r, d, u
            if (mk1[cc - 1]) t = t + ab1[cc - 1];
            if (mk1[cc + 1]) t = t + ab1[cc + 1];
            if (mk2[cc]) t = t + ab2[cc];
            if (mk0[cc]) t = t + ab0[cc];
          }
          const double q = t * pc1[k];
          q0[k] = q;
          if (black) ab1[cc] = q * pc1[k];                         // what a red neighbour adds (times its pc)
        }
      }
    }
    __syncthreads();

    // ---- stage 3: z(yy-2), z.r' -----------------------------------------------------------
    const int y2r = yy - 2;
    if (is_main && ncol > 0 && y2r >= pz.y0 && y2r < pz.y1 && yy >= pz.y0 + 2) {
      DV<C> out;
      bool any = false;
#pragma unroll
      for (int k = 0; k < C; ++k) {
        out.v[k] = 0.0;
        const int cc = col + k;
        if (!mk2[cc]) continue;
        any = true;
        const double pk = pc2[k];
        double zc;
        if ((gx + k + gyy - 2) & 1) {
          zc = q2[k] * pk;                                         // black: q*pc
        } else {
          double t = q2[k];
          if (mk2[cc - 1]) t = t + pk * ab2[cc - 1];
          if (mk2[cc + 1]) t = t + pk * ab2[cc + 1];
          if (mk3[cc]) t = t + pk * ab3[cc];
          if (mk1[cc]) t = t + pk * ab1[cc];
          zc = t * pk;
        }
        out.v[k] = zc;
        if (y2r >= acc0 && y2r < acc1) acc += zc * r2[k];
      }
      // only the owned rows are stored: the halo rows of z belong to the neighbouring slabs
      if (any && y2r >= acc0 && y2r < acc1) {
        const size_t c2 = gidx(g, pz.x0 + t0, y2r);
        stv<C>(z + c2, out);
        if (z_dn && y2r < acc0 + dist.depth) { stv<C>(z_dn + c2, out); peer_stored = true; }
        if (z_up && y2r >= acc1 - dist.depth) { stv<C>(z_up + c2, out); peer_stored = true; }
      }
    }
    // shift the register window
#pragma unroll
    for (int k = 0; k < C; ++k) { r2[k] = r1[k]; pc2[k] = pc1[k]; q2[k] = q0[k]; r1[k] = r0[k]; pc1[k] = pc0[k]; }
    (void)q1;
    cons.next(g, T, th, active.list);
  }

  if (peer_stored) __threadfence_system();
  const double bsum = block_reduce<false>(acc);
  const double bmax = block_reduce<true>(mx);
  // one ticket, two partials per block: sum in [0, B), max in [B, 2B)
  __shared__ bool last_sh;
  __shared__ double tot_sh[2];
  const unsigned int nblocks = gridDim.x;
  if (tid == 0) {
    partials[blockIdx.x] = bsum;
    partials[nblocks + blockIdx.x] = bmax;
    __threadfence();
    last_sh = atomicAdd(&sc->ctr[CTR_ZR], 1u) == nblocks - 1;
  }
  __syncthreads();
  if (!last_sh) return;
  __threadfence();
  double a = 0.0, m = 0.0;
  for (unsigned int i = tid; i < nblocks; i += blockDim.x) {
    a += __ldcg(partials + i);
    m = fmax(m, __ldcg(partials + nblocks + i));
  }
  a = block_reduce<false>(a);
  m = block_reduce<true>(m);
  if (tid == 0) { sc->ctr[CTR_ZR] = 0; tot_sh[0] = a; tot_sh[1] = m; }
  __syncthreads();
  const double total = tot_sh[0], norm = tot_sh[1];
  if (dist.mine) {                                           // halo flags + {z.r, ||r||inf} over NVLink
    if (tid == 0) sc->part[1] = norm;
    __syncthreads();
    p2p_finish(dist, sc, 1, 0, tol, total, true);
    return;
  }
  if (tid != 0) return;
  if (exact == 2) { sc->part[0] = total; sc->part[1] = norm; return; }   // slab mode over NCCL: folded across ranks later
  sc->resid = norm;
  sc->iters += 1;
  if (norm <= tol) { sc->done = 1; return; }                 // main.c:756-758
  sc->beta = total / sc->sigma;                              // main.c:762-765
  sc->sigma = total;
}
