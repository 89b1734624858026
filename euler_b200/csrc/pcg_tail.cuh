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
//     r', pc and q of the two previous rows stay in registers, and so do the fluid masks of its
//     cells and their neighbours (one bit word per row);
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
#include <type_traits>

namespace tail {

constexpr int HALO = 2;                 // halo columns of the value planes (== pipe::Elem<double>::HX)
constexpr int ABW = TW + 2 * HALO;      // doubles per row of the rolling neighbour buffer
#ifndef EULER_TAIL_PF_ROWS
#define EULER_TAIL_PF_ROWS -1
#endif
constexpr int PF_ROWS = EULER_TAIL_PF_ROWS;   // rows ahead of the ring's newest row the p-update operands are prefetched to L2
constexpr int NSLOT = 6;                // steps j-4 .. j are live; a fast warp may already write step j+1

template <int NS>
constexpr int smem_bytes() {
  return NS * pipe::Layout<3, 1>::stage_bytes + 2 * NS * 8 + 2 * NSLOT * ABW * 8;
}

}  // namespace tail

// Per-thread state of the row pipeline, NC cells per thread (C for the main threads, 1 for a lane
// of the halo warp).  Masks travel as bit words: bit i of mw* = cell (col - 1 + i) is fluid, one
// word per row, shifted along with the register window.
template <int NC>
struct TailLane {
  double pc1[NC];                                     // pc of row yy-1
  double pc2[NC], q2[NC], pc3[NC], q3[NC];            // pc, q of rows yy-2 / yy-3
  unsigned mw0, mw1, mw2, mw3, mw4;                   // rows yy .. yy-4
};

__device__ __forceinline__ bool tbit(unsigned w, int i) { return (w >> i) & 1u; }

// mode: 0 = r only (odd iterations), 1 = r and both pending p updates (even iterations),
//       2 = r and this iteration's p update (every iteration; the parity hook)
template <int NS, int C, int MB = 3>
__global__ void __launch_bounds__(TW / C + 32, MB) k_fused_tail(
    Grid g, TileList active, const double* __restrict__ r, const double* __restrict__ as,
    const double* __restrict__ precon, const uint8_t* __restrict__ fluid, const double* __restrict__ s,
    const double* __restrict__ s_prev, double* __restrict__ p, double* __restrict__ r_new,
    double* __restrict__ z, double* partials, DevScalars* sc, double tol, int mode, int exact,
    int acc0, int acc1, const __grid_constant__ DistArgs dist, int split_it, unsigned long long* tr) {
  using L = pipe::Layout<3, 1>;
  constexpr int NMAIN = TW / C;                  // main threads: C cells each
  constexpr int ROWT = L::ROWT;
  if (sc->done) return;
  trace_mark(tr, 0);
  if (tr && blockIdx.x == 0 && threadIdx.x == 0) { tr[8] = 2; tr[9] = (unsigned long long)split_it; }
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* stages = smem_raw;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + NS * L::stage_bytes);
  uint64_t* empty = full + NS;
  double* ab = reinterpret_cast<double*>(empty + NS) + tail::HALO;          // [NSLOT][ABW], index slot*ABW + col
  double* rb = ab + tail::NSLOT * tail::ABW;     // r' of the same rows (stage 2 and 3 read their own cells back:
                                                 // three rows of r' in registers would not fit next to the rest)

  const int tid = threadIdx.x, lane = tid & 31;
  const bool is_main = tid < NMAIN;
  const bool producer = tid == NMAIN;            // lane 0 of the extra warp runs the ring
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { pipe::mbar_init(full + i, 1); pipe::mbar_init(empty + i, NMAIN / 32 + 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int th = g.th;
  const size_t pitch = (size_t)g.pitch;
  const pipe::Tiles T = pipe::tiles_of(g, th);
  pipe::JobIter cons;
  cons.rad = 2;
  cons.start(g, T, th, active.list, (int)*active.count);
  // the producer's own walk over the same rows (NS-1 rows ahead) lives in shared memory: one lane
  // uses it, but registers would be reserved in all 288 threads
  __shared__ pipe::JobIter prod_sh;
  __shared__ int pstage_sh, pphase_sh;           // next stage to fill, parity its `empty` must have passed
  if (producer) { prod_sh = cons; pstage_sh = 0; pphase_sh = 1; }
  bool more_rows = cons.valid;                   // (a register: only the producer lane may touch prod_sh)

  auto issue = [&]() {                           // producer lane only
    pipe::JobIter prod = prod_sh;
    int pstage = pstage_sh, pphase = pphase_sh;
    pipe::mbar_wait(empty + pstage, (uint32_t)pphase);
    unsigned char* dst = stages + pstage * L::stage_bytes;
    const uint32_t b8 = (uint32_t)(prod.p.w + 2 * tail::HALO) * 8u, b1 = (uint32_t)(prod.p.w + 2 * pipe::HX1);
    pipe::mbar_expect_tx(full + pstage, 3 * b8 + b1);
    const long row = (long)prod.yy * g.pitch + prod.p.x0;
    pipe::bulk_g2s(dst, r + row - tail::HALO, b8, full + pstage);
    pipe::bulk_g2s(dst + ROWT, as + row - tail::HALO, b8, full + pstage);
    pipe::bulk_g2s(dst + 2 * ROWT, precon + row - tail::HALO, b8, full + pstage);
    pipe::bulk_g2s(dst + 3 * ROWT, fluid + row - pipe::HX1, b1, full + pstage);
    // the element-wise operands of the p update are read with plain loads when the row is consumed:
    // pull them into L2 now (even iterations, owned rows only)
    // pull them into L2 PF rows before that (even iterations, owned rows only; the first PF rows of
    // a piece go out with its first halo row)
    if (mode != 0) {
      constexpr int PF = tail::PF_ROWS;
      const uint32_t bp = (uint32_t)prod.p.w * 8u;
      auto pf = [&](int y) {
        const long o = (long)y * g.pitch + prod.p.x0;
        pipe::bulk_prefetch_l2(s + o, bp);
        pipe::bulk_prefetch_l2(p + o, bp);
        if (mode == 1) pipe::bulk_prefetch_l2(s_prev + o, bp);
      };
      if (PF > 2 && prod.yy == prod.p.y0 - 2)
        for (int y = prod.p.y0; y < prod.p.y0 + PF - 2 && y < prod.p.y1; ++y) pf(y);
      const int t = prod.yy + PF;
      if (t >= prod.p.y0 + (PF > 2 ? PF - 2 : 0) && t < prod.p.y1) pf(t);
    }
    if (++pstage == NS) { pstage = 0; pphase ^= 1; }
    prod.next(g, T, th, active.list);
    prod_sh = prod; pstage_sh = pstage; pphase_sh = pphase;
    more_rows = prod.valid;
  };
  if (producer)
    for (int i = 0; i < NS - 1 && more_rows; ++i) issue();       // NS-1 rows ahead from here on

  double alpha = sc->alpha, alpha_prev = sc->alpha_prev;
  if (split_it) {
    // split-phase exchange (p2p.cuh): every block consumes the {z.s} partials the search/apply
    // kernel posted and forms alpha itself (main.c:752); block 0 records it.  The first rows of the
    // ring are already in flight: none of this kernel's inputs comes from another rank.
    double zs, unused;
    const bool ok = p2p_collect(dist, false, zs, unused);
    const bool writer = blockIdx.x == 0 && threadIdx.x == 0;
    // a peer that never showed up (bounded poll): flag it — every later kernel returns at once — but
    // finish this launch normally (bulk copies are in flight into this block's shared memory)
    if (!ok && writer) { sc->comm_timeout = 1; sc->done = 1; }
    alpha = sc->sigma_s[(split_it + 1) & 1] / zs;
    alpha_prev = sc->alpha_s[(split_it + 1) & 1];          // the previous iteration's
    if (writer) { sc->alpha_s[split_it & 1] = alpha; sc->zs = zs; sc->alpha_prev = alpha_prev; sc->alpha = alpha; }
  }
  trace_mark(tr, 1);
  trace_block(tr, sc, 2, 0);
  const double neg_alpha = -alpha;
  // alpha and alpha_prev are needed by the p update only (even iterations): block-wide constants kept
  // in shared memory rather than in four registers of every thread
  __shared__ double alpha_sh[2];
  if (tid == 0) { alpha_sh[0] = alpha; alpha_sh[1] = alpha_prev; }
  __syncthreads();

  TailLane<C> t;
#pragma unroll
  for (int k = 0; k < C; ++k) t.pc1[k] = t.pc2[k] = t.q2[k] = t.pc3[k] = t.q3[k] = 0.0;
  t.mw0 = t.mw1 = t.mw2 = t.mw3 = t.mw4 = 0u;
  double acc = 0.0, mx = 0.0;
  bool peer_stored = false;
  double* __restrict__ z_dn = dist.z_dn;
  double* __restrict__ z_up = dist.z_up;
  const bool has_peer = z_dn || z_up;
  // rolling row slots of the neighbour buffer, as element offsets: steps j .. j-4 and the free one.
  // ONE barrier per step: stage 2 works on row yy-1 (needs the red entries of rows yy-2 .. yy, the
  // last written by this step's stage 1, hence the barrier), stage 3 on row yy-3 (needs the black
  // entries of rows yy-4 .. yy-2, all written by EARLIER steps' stage 2), so the two run back to
  // back without a barrier between them; with 6 slots the one a fast warp already overwrites in
  // the next step's stage 1 (step j-5's) is one nobody still reads.  The last row of a piece is
  // finished by a second stage-3 pass in the piece's last step, behind one extra barrier.
  int o0 = 0, o1 = tail::ABW, o2 = 2 * tail::ABW, o3 = 3 * tail::ABW, o4 = 4 * tail::ABW, of = 5 * tail::ABW;
  int cstage = 0, cphase = 0;                    // consumer: ring stage of this step, its parity
  // per piece: this thread's first column / cell count, global parity base, element offset of
  // (x0 + col, yy) in the planes
  int col = 0, ncol = 0, par_base = 0;
  size_t rowp = 0;

  while (cons.valid) {
    if (producer && more_rows) issue();
    const pipe::Piece& pz = cons.p;
    const int yy = cons.yy;                      // row whose inputs arrive in this step
    const int rel = yy - pz.y0;                  // -2 at the first row of a piece
    if (rel == -2) {
      if (is_main) { col = tid * C; ncol = col < pz.w ? C : 0; }
      else { col = lane < 2 ? lane - 2 : pz.w + lane - 2; ncol = lane < 4 ? 1 : 0; }
      par_base = pz.x0 + col + g.yoff;
      rowp = (size_t)yy * pitch + (size_t)(pz.x0 + col);
    }
    const int par0 = (par_base + yy) & 1;        // parity of (col, yy): 0 = red
    const bool own_row = (unsigned)rel < (unsigned)(pz.y1 - pz.y0);
    const bool acc_row = yy >= acc0 && yy < acc1;
    const bool last_step = yy == pz.y1 + 1;      // block-uniform

    // the element-wise operands of the p update, issued before the wait on the ring
    const bool do_p = mode != 0 && own_row && is_main && ncol > 0;
    DV<C> sv, spv, pv;
    if (do_p) {
      sv = ldg_v<C>(s + rowp); pv = ldg_v<C>(p + rowp);
      if (mode == 1) spv = ldg_v<C>(s_prev + rowp);
    }

    // ---- stage 1: r'(yy), w(yy) ---------------------------------------------------------
    pipe::mbar_wait(full + cstage, (uint32_t)cphase);
    double r0[C], pc0[C];
    {
      const pipe::RowView<3, 1> in = L::view(stages + cstage * L::stage_bytes);
      unsigned mw = 0u;
      if (is_main) {
        if (ncol) {
          const DV<C> rr = ldsv<C>(in.d[0] + col), aa = ldsv<C>(in.d[1] + col), pp = ldsv<C>(in.d[2] + col);
#pragma unroll
          for (int i = 0; i < C + 2; ++i) mw |= (in.b[0][col - 1 + i] ? 1u : 0u) << i;
          // (main threads start on an even column of a 512-aligned piece: the colour of their cell k
          // in this row is the same for the whole block, so the two colourings are two straight-line
          // instantiations instead of per-cell predicates)
          auto red_w = [&](auto parc) {
            constexpr int PAR = decltype(parc)::value;
#pragma unroll
            for (int k = 0; k < C; ++k)
              if (((PAR + k) & 1) == 0) ab[o0 + col + k] = pp.v[k] * (r0[k] * pp.v[k]);   // red: what a black neighbour adds
          };
#pragma unroll
          for (int k = 0; k < C; ++k) {
            const bool m = tbit(mw, k + 1);
            const double rn = m ? rr.v[k] + aa.v[k] * neg_alpha : rr.v[k];      // fmadd(z, -alpha, r), main.c:754
            r0[k] = rn; pc0[k] = pp.v[k];
            if (own_row && m) {
              if (mode == 1) pv.v[k] = pv.v[k] + spv.v[k] * alpha_sh[1];        // the previous iteration's main.c:753
              if (mode) pv.v[k] = pv.v[k] + sv.v[k] * alpha_sh[0];              // fmadd(s, alpha, p), main.c:753
              // NaN-dropping max like main.c:659-662 (`if (a > max) max = a`): fmax returns the other
              // operand for a NaN, and |r'| >= +0 so the sign of a zero cannot matter
              if (acc_row) mx = fmax(mx, fabs(rn));
            }
          }
          if (par0) red_w(std::integral_constant<int, 1>{}); else red_w(std::integral_constant<int, 0>{});
          {
            DV<C> o;
#pragma unroll
            for (int k = 0; k < C; ++k) o.v[k] = r0[k];
            stv<C>(rb + o0 + col, o);
            if (own_row) {
              stv<C>(r_new + rowp, o);
              if (do_p) stv<C>(p + rowp, pv);
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < C; ++k) r0[k] = pc0[k] = 0.0;
        }
      } else {
#pragma unroll
        for (int k = 0; k < C; ++k) r0[k] = pc0[k] = 0.0;
        if (ncol) {                              // one halo column per lane
#pragma unroll
          for (int i = 0; i < 3; ++i) mw |= (in.b[0][col - 1 + i] ? 1u : 0u) << i;
          const double rr = in.d[0][col], aa = in.d[1][col], pp = in.d[2][col];
          const double rn = tbit(mw, 1) ? rr + aa * neg_alpha : rr;
          r0[0] = rn; pc0[0] = pp;
          rb[o0 + col] = rn;
          if (par0 == 0) ab[o0 + col] = pp * (rn * pp);
        }
      }
      t.mw0 = mw;
      // the ring stage is free again
      __syncwarp();
      if (lane == 0) pipe::mbar_arrive(empty + cstage);
      if (++cstage == NS) { cstage = 0; cphase ^= 1; }
    }
    __syncthreads();

    // ---- stage 2: q(yy-1), zb(yy-1): rows y0-1 .. y1 ----------------------------------------
    double q1[C];
#pragma unroll
    for (int k = 0; k < C; ++k) q1[k] = 0.0;
    if (rel >= 0 && ncol) {
      if (is_main) {
        const DV<C> r1 = ldsv<C>(rb + o1 + col);
        auto fwd = [&](auto parc) {
          constexpr int PAR = decltype(parc)::value;
#pragma unroll
          for (int k = 0; k < C; ++k) {
            if (!tbit(t.mw1, k + 1)) continue;
            const int cc = col + k;
            double v = r1.v[k];
            const bool black = ((PAR + k) & 1) == 0;                 // row yy-1: the colours of row yy swapped (folds after unrolling)
            if (black) {                                             // + red neighbours: l, r, d, u
              if (tbit(t.mw1, k)) v = v + ab[o1 + cc - 1];
              if (tbit(t.mw1, k + 2)) v = v + ab[o1 + cc + 1];
              if (tbit(t.mw2, k + 1)) v = v + ab[o2 + cc];
              if (tbit(t.mw0, k + 1)) v = v + ab[o0 + cc];
            }
            const double q = v * t.pc1[k];
            q1[k] = q;
            if (black) ab[o1 + cc] = q * t.pc1[k];                   // what a red neighbour adds (times its pc)
          }
        };
        if (par0) fwd(std::integral_constant<int, 1>{}); else fwd(std::integral_constant<int, 0>{});
      } else if ((lane == 1 || lane == 2) && tbit(t.mw1, 1) && par0 == 0) {
        // inner halo columns -1 and w: only a black cell's zb is ever asked for
        double v = rb[o1 + col];
        if (tbit(t.mw1, 0)) v = v + ab[o1 + col - 1];
        if (tbit(t.mw1, 2)) v = v + ab[o1 + col + 1];
        if (tbit(t.mw2, 1)) v = v + ab[o2 + col];
        if (tbit(t.mw0, 1)) v = v + ab[o0 + col];
        ab[o1 + col] = (v * t.pc1[0]) * t.pc1[0];
      }
    }

    // ---- stage 3: z(row), z.r' for a row whose neighbours' zb are complete ------------------
    // `back` = how many rows behind yy (3, or 2 in the second pass of a piece's last step); the
    // row's own q / pc / r', its mask words (centre, below, above) and buffer rows come with it
    auto stage3 = [&](int back, const double (&qv)[C], const double (&pcv)[C], unsigned mc,
                      unsigned mdn, unsigned mup, int oc, int odn, int oup) {
      const int yrow = yy - back;
      const bool accr = yrow >= acc0 && yrow < acc1;
      const DV<C> rv = ldsv<C>(rb + oc + col);
      DV<C> out;
      const unsigned any = mc & (((1u << C) - 1u) << 1);
      auto bwd = [&](auto parc) {
        constexpr int PAR = decltype(parc)::value;   // 1: cell 0 of this row is black
#pragma unroll
        for (int k = 0; k < C; ++k) {
          out.v[k] = 0.0;
          if (!tbit(mc, k + 1)) continue;
          const int cc = col + k;
          const double pk = pcv[k];
          double zc;
          if (((PAR + k) & 1) != 0) {
            zc = qv[k] * pk;                                         // black: q*pc
          } else {
            double v = qv[k];
            if (tbit(mc, k)) v = v + pk * ab[oc + cc - 1];
            if (tbit(mc, k + 2)) v = v + pk * ab[oc + cc + 1];
            if (tbit(mdn, k + 1)) v = v + pk * ab[odn + cc];
            if (tbit(mup, k + 1)) v = v + pk * ab[oup + cc];
            zc = v * pk;
          }
          out.v[k] = zc;
          if (accr) acc += zc * rv.v[k];
        }
      };
      // colour of cell 0 in row yy - back: that of row yy when back is even
      if (((par0 + back) & 1) != 0) bwd(std::integral_constant<int, 1>{}); else bwd(std::integral_constant<int, 0>{});
      // only the owned rows are stored: the halo rows of z belong to the neighbouring slabs
      if (any && accr) {
        const size_t c2 = rowp - (size_t)back * pitch;
        stv<C>(z + c2, out);
        if (has_peer) {
          if (z_dn && yrow < acc0 + dist.depth) { stv<C>(z_dn + c2, out); peer_stored = true; }
          if (z_up && yrow >= acc1 - dist.depth) { stv<C>(z_up + c2, out); peer_stored = true; }
        }
      }
    };
    const bool main_live = is_main && ncol > 0;
    if (main_live && rel >= 3) stage3(3, t.q3, t.pc3, t.mw3, t.mw4, t.mw2, o3, o4, o2);
    if (last_step) {
      // the piece ends here: its last row (yy-2) still needs the zb this step's stage 2 just wrote
      __syncthreads();
      if (main_live && rel >= 2) stage3(2, t.q2, t.pc2, t.mw2, t.mw3, t.mw1, o2, o3, o1);
    }
    // shift the register window and the row slots
#pragma unroll
    for (int k = 0; k < C; ++k) {
      t.pc3[k] = t.pc2[k]; t.q3[k] = t.q2[k];
      t.pc2[k] = t.pc1[k]; t.q2[k] = q1[k]; t.pc1[k] = pc0[k];
    }
    t.mw4 = t.mw3; t.mw3 = t.mw2; t.mw2 = t.mw1; t.mw1 = t.mw0;
    { const int f = o4; o4 = o3; o3 = o2; o2 = o1; o1 = o0; o0 = of; of = f; }
    rowp += pitch;
    cons.next(g, T, th, active.list);
  }

  trace_mark(tr, 2);
  trace_block(tr, sc, 2, 1);
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
  if (!last_sh) { trace_mark(tr, 3); return; }
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
    if (split_it) { p2p_post(dist, total, norm, true); trace_mark(tr, 3); return; }
    if (tid == 0) sc->part[1] = norm;
    __syncthreads();
    p2p_finish(dist, sc, 1, 0, tol, total, true);
    trace_mark(tr, 3);
    return;
  }
  if (tid != 0) return;
  if (exact == 2) { sc->part[0] = total; sc->part[1] = norm; return; }   // slab mode over NCCL: folded across ranks later
  sc->resid = norm;
  sc->iters += 1;
  if (norm <= tol) { sc->done = 1; return; }                 // main.c:756-758
  sc->beta = total / sc->sigma;                              // main.c:762-765
  sc->sigma = total;
  trace_mark(tr, 3);
}
