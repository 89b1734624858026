// euler_b200/csrc/pcg_kernels.cu — Bridson's PCG for the pressure Poisson system, the part of
// project() at reference main.c:735-767, with the helpers it calls:
//
//   k_apply_a        z = A s, fused with the z.s reduction and alpha      main.c:679-691, 752
//   k_axpy           p += alpha s ; r -= alpha z ; ||r||inf ; stop test   main.c:694-702, 753-758
//   k_update_search  s = z + beta s                                       main.c:669-677, 764
//   k_rb_*           red-black ordered IC(0) preconditioner (GPU-parallel mode; not in the
//                    reference — CPU mirror in oracle/euler_oracle.c precon_redblack)
//
// The reference-faithful natural-order IC(0) lives in wavefront.cu.
//
// Execution shape (all kernels here): a PERSISTENT grid of sm_count x PCG_BLOCKS_PER_SM blocks
// walks the grid in tiles of TW x TH = 512 x 32 cells.  A thread owns 4 consecutive x (one
// uchar4 / two double2 loads per plane and row, 16 B-aligned, fully coalesced) and marches
// up the TH rows of the tile keeping a three-row window of the stencil operand in
// REGISTERS, so every operand row is loaded from HBM/L2 once per tile (plus one halo row
// above and below).  Left/right neighbours come from the adjacent lane by warp shuffle; only
// the two edge lanes of a warp issue one extra scalar load.  Tiles without any fluid cell are
// skipped through a per-tile flag built once per sub-step (k_tile_flags) — like the
// reference, which touches only is_fluid cells.
//
// Scalars never visit the host: each reducing kernel leaves ONE partial per block, the block
// that finishes last folds them in a fixed order (deterministic) and writes alpha / beta /
// sigma / the stop flag into DevScalars.  Once `done` is set every later kernel of the solve
// returns immediately, so the host may enqueue iterations in batches and poll rarely; the
// iterate sequence is the same as with the reference's per-iteration test.
//
// All vectors are fp64 like the reference's.  -fmad=false + source-order arithmetic make
// every element-wise result bit-identical to the CPU's; only the order of the dot-product
// sums differs.
//
// Mixed-precision mode (euler_params.pcg_dtype = EULER_PCG_FP32; not in the
// reference): the kernels of the fused red-black iteration are templates on the STORAGE type T
// of r, z, s, q, A s and the preconditioner diagonal.  With T = float every element-wise
// operation is one fp32 operation in the same order (CPU mirror: oracle/euler_oracle.c
// pcg_mixed), p stays fp64, the dot products multiply and accumulate in fp64, and every
// `pcg_refresh_every` iterations k_true_residual replaces r by b - A p evaluated in fp64.
// 73 B/cell per iteration instead of 132.  T = double instantiates exactly the code that was
// here before the template (checked: identical SASS).
#include "kernels.h"
#include "p2p.cuh"
#include "pcg_pipe.cuh"
#include "pcg_ops.cuh"

#include <stdlib.h>
#include <string.h>

#include <map>
#include <utility>

namespace euler {

namespace {

constexpr int TW = 512, TH = 32, TT = 128;     // tile width/height in cells, threads per block
enum { CTR_ZS = 0, CTR_NORM = 1, CTR_ZR = 2 };

struct Tiles { int tx, ty, n; };
__host__ __device__ inline Tiles tiles_of(const Grid& g) {
  Tiles t;
  t.tx = (g.pitch + TW - 1) / TW;
  t.ty = (g.ny + g.th - 1) / g.th;
  t.n = t.tx * t.ty;
  return t;
}

struct TileList { const int* __restrict__ list; const unsigned int* __restrict__ count; };


// ---- programmatic dependent launch (PDL): tried, measured, OFF by default -------------------
// The four kernels of a PCG iteration can be launched with the programmatic-stream-
// serialisation attribute (EULER_PDL=1): their blocks may then become resident while the
// previous kernel is still in its tail and park at `griddepcontrol.wait`, which returns once
// the previous kernel has completed and its writes are visible; nothing of the previous
// kernel's output is touched before that wait, so results are unchanged (GPU suite green both
// ways).  The hope was to hide launch latency and prologue between dependent kernels.
// Measured on B200 at 16384^2 (same box, back to back, no per-launch timers): 167.6 ms per
// sub-step with PDL against 144.2 ms without, and 100.1 against 82.5 ms on 2 slabs — every
// programmatic boundary costs ~55 us more than a plain stream-ordered one here, so the plain
// launch stays the default.  With EULER_PDL unset `griddepcontrol.*` are no-ops.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

static bool pdl_enabled() {
  static const int v = getenv("EULER_PDL") ? atoi(getenv("EULER_PDL")) : 0;
  return v != 0;
}

template <class... KArgs, class... Args>
static void launch_pdl(void (*kernel)(KArgs...), int blocks, int threads, size_t smem, cudaStream_t stream,
                       Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}


// per-tile "contains fluid" flags
__global__ void __launch_bounds__(TT) k_tile_flags(Grid g, const uint8_t* __restrict__ fluid,
                                                   uint8_t* __restrict__ active, DevScalars* sc) {
  const Tiles T = tiles_of(g);
  for (int tile = blockIdx.x; tile < T.n; tile += gridDim.x) {
    const int x0 = (tile % T.tx) * TW + threadIdx.x * 4, y0 = (tile / T.tx) * g.th;
    const int y1 = min(y0 + g.th, g.ny);
    unsigned any = 0;
    if (x0 < g.pitch)
      for (int y = y0; y < y1; ++y) any |= ldmask(fluid + gidx(g, x0, y));
    const int has = __syncthreads_or(any != 0);
    if (threadIdx.x == 0) active[tile] = has ? 1 : 0;
  }
}

// Ordered, compact list of the tiles that contain fluid (single block): the persistent
// kernels walk list[blockIdx.x + i*gridDim.x], which balances the blocks to within one tile
// whatever the shape of the fluid region.
__global__ void __launch_bounds__(1024) k_tile_compact(Grid g, const uint8_t* __restrict__ active,
                                                       int* __restrict__ list, DevScalars* sc) {
  const Tiles T = tiles_of(g);
  __shared__ int sh[1024];
  const int per = (T.n + 1023) / 1024;
  const int lo = min(T.n, per * (int)threadIdx.x), hi = min(T.n, lo + per);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) cnt += active[i] ? 1 : 0;
  sh[threadIdx.x] = cnt;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const int v = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
    __syncthreads();
    sh[threadIdx.x] += v;
    __syncthreads();
  }
  int run = sh[threadIdx.x] - cnt;
  for (int i = lo; i < hi; ++i)
    if (active[i]) list[run++] = i;
  if (threadIdx.x == 1023) sc->active_tiles = (unsigned int)sh[1023];
}

// ---------------------------------------------------------------------------------------
// Three-row register window over a per-cell operand W (what the 5-point stencil reads from
// the neighbours) plus the fluid mask.  `Op` supplies:
//   D4   Op::operand(c)            W of the 4 cells starting at flat index c
//   double Op::operand1(c)         W of the single cell c (edge lanes)
//   void Op::cell(c, k, m, wc, wl, fl, wr, fr, wd, fd, wu, fu)   consume one cell
//   operand(c, slot) is called once per row entering the window (slot says where), so an Op
//   may stash other planes of that row; rotate() is called when UP becomes CENTER.
// Call order inside a tile row is k = 0..3 for ascending x.
enum { SLOT_DOWN = 0, SLOT_CENTER = 1, SLOT_UP = 2 };

template <class Op>
__device__ __forceinline__ void stencil_tile(const Grid& g, const uint8_t* __restrict__ fluid,
                                             int x0, int y0, int y1, bool live, Op& op) {
  const int lane = threadIdx.x & 31;
  const unsigned keep = live ? 0xffffffffu : 0u;   // threads past the row end present no fluid
  size_t c = gidx(g, x0, y0);
  D4 wd = op.operand(c - g.pitch, SLOT_DOWN);
  unsigned md = ldmask(fluid + c - g.pitch) & keep;
  D4 wc = op.operand(c, SLOT_CENTER);
  unsigned mc = ldmask(fluid + c) & keep;
  for (int y = y0; y < y1; ++y, c += g.pitch) {
    const D4 wu = op.operand(c + g.pitch, SLOT_UP);
    const unsigned mu = ldmask(fluid + c + g.pitch) & keep;
    // horizontal neighbours of the quad's two end cells
    double wl = __shfl_up_sync(EULER_FULL_MASK, wc.v[3], 1);
    double wr = __shfl_down_sync(EULER_FULL_MASK, wc.v[0], 1);
    unsigned ml = __shfl_up_sync(EULER_FULL_MASK, mc, 1) >> 24;
    unsigned mr = __shfl_down_sync(EULER_FULL_MASK, mc, 1) & 0xffu;
    if (lane == 0) { ml = fluid[c - 1]; wl = ml ? op.operand1(c - 1) : 0.0; }
    if (lane == 31) { mr = fluid[c + 4]; wr = mr ? op.operand1(c + 4) : 0.0; }
    if (mc) {
      op.begin_row(c, mc);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!mbit(mc, k)) continue;
        const double l = k == 0 ? wl : wc.v[k - 1];
        const bool fl = k == 0 ? (ml != 0) : mbit(mc, k - 1);
        const double r = k == 3 ? wr : wc.v[k + 1];
        const bool fr = k == 3 ? (mr != 0) : mbit(mc, k + 1);
        op.cell(c, k, wc.v[k], l, fl, r, fr, wd.v[k], mbit(md, k), wu.v[k], mbit(mu, k), x0 + k, y + g.yoff);
      }
      op.end_row(c, mc);
    }
    wd = wc; md = mc; wc = wu; mc = mu;
    op.rotate();
  }
}

template <class Body>
__device__ __forceinline__ void for_each_tile(const Grid& g, const TileList& tl, Body body) {
  // balanced split of the active tiles' rows over the persistent grid (pcg_pipe.cuh ChunkIter)
  const pipe::Tiles T = pipe::tiles_of(g, g.th);
  pipe::ChunkIter ch;
  ch.start((int)*tl.count, g.th);
  pipe::Piece p;
  while (ch.next(g, T, g.th, tl.list, p)) {
    const int x0 = p.x0 + threadIdx.x * 4;
    // threads past the row end keep participating in the shuffles with a clamped, harmless
    // address (their mask is the zero padding / they are never asked for a valid neighbour)
    const int xs = min(x0, g.pitch - 4);
    body(xs, p.y0, p.y1, x0 < g.pitch);
  }
}

// ---- z = A s --------------------------------------------------------------------------
struct ApplyA {
  const double* __restrict__ s;
  const int8_t* __restrict__ adiag;
  double* __restrict__ z;
  double acc;
  int a0, a1;                      // rows whose cells count towards the dot product
  unsigned am;
  D4 out;
  __device__ __forceinline__ D4 operand(size_t c, int) const { return ld4(s + c); }
  __device__ __forceinline__ double operand1(size_t c) const { return s[c]; }
  __device__ __forceinline__ void rotate() {}
  __device__ __forceinline__ void begin_row(size_t c, unsigned) {
    am = ldmask(reinterpret_cast<const uint8_t*>(adiag) + c);
    out.v[0] = out.v[1] = out.v[2] = out.v[3] = 0.0;
  }
  __device__ __forceinline__ void cell(size_t, int k, double sc, double l, bool fl, double r, bool fr,
                                       double d, bool fd, double u, bool fu, int, int gy) {
    // main.c:683-687: a_diag*s - right - up - left - down, each only towards fluid
    double o = (double)(int)(signed char)((am >> (8 * k)) & 0xffu) * sc;
    o -= fr ? r : 0.0;
    o -= fu ? u : 0.0;
    o -= fl ? l : 0.0;
    o -= fd ? d : 0.0;
    out.v[k] = o;
    if (gy >= a0 && gy < a1) acc += o * sc;
  }
  __device__ __forceinline__ void end_row(size_t c, unsigned) { st4(z + c, out); }
};

__global__ void __launch_bounds__(TT) k_apply_a(
    Grid g, TileList active, const double* __restrict__ s,
    const uint8_t* __restrict__ fluid, const int8_t* __restrict__ adiag, double* __restrict__ z,
    double* partials, DevScalars* sc, int exact, int acc0, int acc1) {
  if (sc->done) return;
  ApplyA op{s, adiag, z, 0.0, acc0 + g.yoff, acc1 + g.yoff, 0u, {}};
  for_each_tile(g, active, [&](int x0, int y0, int y1, bool live) {
    stencil_tile(g, fluid, x0, y0, y1, live, op);
  });
  const double bsum = block_reduce<false>(op.acc);
  grid_reduce_last_block<false>(bsum, partials, &sc->ctr[CTR_ZS], [&](double total) {
    if (exact == 2) { sc->part[0] = total; return; }         // slab mode: summed over ranks later
    if (exact) return;                                       // k_dot_seq supplies z.s instead
    sc->zs = total;
    sc->alpha_prev = sc->alpha; sc->alpha = sc->sigma / total;                           // main.c:752
  });
}

// ---- p += alpha s ; r -= alpha z ; ||r||inf ---------------------------------------------
// `mode` (fused red-black iteration only; 2 = both updates every iteration, as the reference):
// the two search directions of consecutive iterations live in two planes (s ping-pongs, see
// k_fused_search_apply), so p can take TWO updates in one pass every second iteration,
//   mode 0 (odd iterations)   r -= alpha (A s)                                  24 B/cell
//   mode 1 (even iterations)  p = (p + alpha' s') + alpha s ; r -= alpha (A s)  56 B/cell
// instead of 48 B/cell every iteration: p is read and written half as often and s is not read
// again on odd iterations.  Same operations in the same order on every cell, so p is bit-
// identical to the per-iteration update (main.c:753); a solve that ends on an odd iteration is
// completed by k_p_fixup.
// T = storage type of s, A s and r (p is fp64 in both modes; with T = float alpha is narrowed
// once for the r update and s is widened for the p update).
// (The per-quad body stays inline on purpose: moved into a function of pcg_ops.cuh like the row
// operators, nvcc 12.9 stopped using the read-only path for these loads and re-ordered the
// memory operations of this, the dominant kernel — checked with cuobjdump, not measured, so not
// taken.  The arithmetic is three fmadds per cell; the GPU suite covers it.)
template <class T>
__global__ void __launch_bounds__(TT) k_axpy(
    Grid g, TileList active, const T* __restrict__ s, const T* __restrict__ s_prev,
    const T* __restrict__ z, const uint8_t* __restrict__ fluid, double* __restrict__ p,
    T* __restrict__ r, double* partials, DevScalars* sc, double tol, int defer, int acc0,
    int acc1, int mode) {
  using V4 = typename Vec4<T>::type;
  pdl_prologue();
  if (sc->done) return;
  const double alpha = sc->alpha, alpha_prev = sc->alpha_prev;
  const T neg_alpha = (T)-alpha;
  T m = (T)0;
  for_each_tile(g, active, [&](int x0, int y0, int y1, bool live) {
    if (!live) return;
    size_t c = gidx(g, x0, y0);
    for (int y = y0; y < y1; ++y, c += g.pitch) {
      const unsigned mc = ldmask(fluid + c);
      if (!mc) continue;
      const V4 zv = ld4(z + c);
      V4 rv = ld4(r + c);
      V4 sv, spv;
      D4 pv;
      if (mode) { sv = ld4(s + c); pv = ld4(p + c); }
      if (mode == 1) spv = ld4(s_prev + c);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!mbit(mc, k)) continue;
        if (mode == 1) pv.v[k] = pv.v[k] + (double)spv.v[k] * alpha_prev;   // the previous iteration's main.c:753
        if (mode) pv.v[k] = pv.v[k] + (double)sv.v[k] * alpha;       // fmadd(s, alpha, p)  main.c:753
        rv.v[k] = rv.v[k] + zv.v[k] * neg_alpha;             // fmadd(z, -alpha, r) main.c:754
        const T a = abs_of(rv.v[k]);
        if (a > m && y >= acc0 && y < acc1) m = a;           // NaN-dropping, main.c:659-662
      }
      if (mode) st4(p + c, pv);
      st4(r + c, rv);
    }
  });
  const double bmax = block_reduce<true>((double)m);
  grid_reduce_last_block<true>(bmax, partials, &sc->ctr[CTR_NORM], [&](double total) {
    if (defer) { sc->part[1] = total; return; }              // slab mode: max over ranks later
    sc->resid = total;
    sc->iters += 1;
    if (total <= tol) sc->done = 1;                          // main.c:756-758
  });
}

// The deferred p update of a solve that stopped after an ODD number of iterations: the last
// iteration's p += alpha s is still pending (s of iteration i lives in plane i & 1).
template <class T>
__global__ void __launch_bounds__(TT) k_p_fixup(Grid g, TileList active, const T* __restrict__ s_odd,
                                                const uint8_t* __restrict__ fluid, double* __restrict__ p,
                                                const DevScalars* sc, int split) {
  using V4 = typename Vec4<T>::type;
  if (!(sc->iters & 1)) return;
  const double alpha = split ? sc->alpha_s[1] : sc->alpha;   // split-phase exchange: alpha of an odd iteration
  for_each_tile(g, active, [&](int x0, int y0, int y1, bool live) {
    if (!live) return;
    size_t c = gidx(g, x0, y0);
    for (int y = y0; y < y1; ++y, c += g.pitch) {
      const unsigned mc = ldmask(fluid + c);
      if (!mc) continue;
      const V4 sv = ld4(s_odd + c);
      D4 pv = ld4(p + c);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (mbit(mc, k)) pv.v[k] = pv.v[k] + (double)sv.v[k] * alpha;
      st4(p + c, pv);
    }
  });
}

// ---- s = z + beta s -------------------------------------------------------------------
__global__ void __launch_bounds__(TT) k_update_search(
    Grid g, TileList active, const double* __restrict__ z,
    const uint8_t* __restrict__ fluid, double* __restrict__ s, const DevScalars* sc) {
  if (sc->done) return;
  const double beta = sc->beta;
  for_each_tile(g, active, [&](int x0, int y0, int y1, bool live) {
    if (!live) return;
    size_t c = gidx(g, x0, y0);
    for (int y = y0; y < y1; ++y, c += g.pitch) {
      const unsigned mc = ldmask(fluid + c);
      if (!mc) continue;
      const D4 zv = ld4(z + c);
      D4 sv = ld4(s + c);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (mbit(mc, k)) sv.v[k] = zv.v[k] + beta * sv.v[k];  // main.c:673
      st4(s + c, sv);
    }
  });
}

// s = z on the active tiles (memcpy(s, z), main.c:746)
__global__ void __launch_bounds__(TT) k_copy_search(Grid g, TileList active,
                                                    const double* __restrict__ z,
                                                    double* __restrict__ s) {
  for_each_tile(g, active, [&](int x0, int y0, int y1, bool live) {
    if (!live) return;
    size_t c = gidx(g, x0, y0);
    for (int y = y0; y < y1; ++y, c += g.pitch) st4(s + c, ld4(z + c));
  });
}

__global__ void k_pcg_reset(DevScalars* sc) {
  sc->iters = 0; sc->done = 0; sc->resid = 0.0; sc->sigma = 0.0; sc->alpha = 0.0; sc->beta = 0.0;
}

// ----------------------------------------------- red-black ordered IC(0) ----
// Red = (x+y) even, eliminated first.  E_red = a (1 if a==0); E_black = a - sum over fluid
// red neighbours of 1/E_red, with the sigma=0.25 safety rule of main.c:594-596.  Neighbour
// sums always run left, right, down, up.  precon = 1/sqrt(E).

__device__ __forceinline__ double rb_e_red(const int8_t* __restrict__ adiag, size_t c) {
  const double a = (double)adiag[c];
  return a != 0.0 ? a : 1.0;
}

// A thread looks at four consecutive cells through one 32-bit mask load and leaves at once
// when none of them is fluid (most of a free-surface scene).
template <class T>
__global__ void __launch_bounds__(256) k_rb_build(
    Grid g, GridTiles gt, const uint8_t* __restrict__ fluid, const int8_t* __restrict__ adiag,
    T* __restrict__ precon) {
  // persistent over the 128-cell x 8-row pieces of the grid-stage tile list (common.cuh GridTiles)
  const unsigned int n = gt.list ? *gt.count : (unsigned int)(gt.tx * gt.ty);
  for (unsigned int i = blockIdx.x; i < n * GT_SUB; i += gridDim.x) {
    const int tile = gt.list ? gt.list[i / GT_SUB] : (int)(i / GT_SUB);
    const int sub = (int)(i % GT_SUB);
    const int x0 = (((tile % gt.tx) * 4 + (sub & 3)) * 32 + threadIdx.x) * 4;
    const int y = ((tile / gt.tx) * 4 + (sub >> 2)) * 8 + threadIdx.y;
    if (x0 >= g.pitch || y >= g.ny || y + g.yoff < 1 || y + g.yoff >= g.gny - 1) continue;
    const unsigned mf = ldmask(fluid + gidx(g, x0, y));
    if (!mf) continue;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + k;
      if (!mbit(mf, k) || x < 1 || x >= g.nx - 1) continue;
      const size_t c = gidx(g, x, y);
      if (((x + y + g.yoff) & 1) == 0) { precon[c] = (T)(1.0 / sqrt(rb_e_red(adiag, c))); continue; }
      const double a = (double)adiag[c];
      double e = a;
      const long off[4] = {-1, 1, -(long)g.pitch, (long)g.pitch};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t nb = c + off[j];
        if (fluid[nb]) e = e - 1.0 / rb_e_red(adiag, nb);
      }
      if (e < 0.25 * a) e = a != 0.0 ? a : 1.0;
      precon[c] = (T)(1.0 / sqrt(e));                        // fp64 factor, narrowed once for T = float
    }
  }
}

// q = L^-1 r.  Red: q = r*pc.  Black: q = (r + sum_nb pc_nb*(r_nb*pc_nb)) * pc.  The window
// operand is W = pc*(r*pc), i.e. pc_nb*q_nb of a red neighbour, recomputed on the fly with
// the same two roundings as the stored q would have.
struct RbForward {
  const double* __restrict__ r;
  const double* __restrict__ pc;
  double* __restrict__ q;
  D4 rr, pp, ru, pu, out;          // r, pc of the centre row / of the row above
  __device__ __forceinline__ D4 operand(size_t c, int slot) {
    const D4 a = ld4(r + c), b = ld4(pc + c);
    if (slot == SLOT_CENTER) { rr = a; pp = b; }
    if (slot == SLOT_UP) { ru = a; pu = b; }
    D4 w;
#pragma unroll
    for (int k = 0; k < 4; ++k) w.v[k] = b.v[k] * (a.v[k] * b.v[k]);
    return w;
  }
  __device__ __forceinline__ double operand1(size_t c) const { const double b = pc[c]; return b * (r[c] * b); }
  __device__ __forceinline__ void rotate() { rr = ru; pp = pu; }
  __device__ __forceinline__ void begin_row(size_t, unsigned) {
    out.v[0] = out.v[1] = out.v[2] = out.v[3] = 0.0;
  }
  __device__ __forceinline__ void cell(size_t, int k, double, double l, bool fl, double rt, bool fr,
                                       double d, bool fd, double u, bool fu, int x, int y) {
    double t = rr.v[k];
    if ((x + y) & 1) {
      if (fl) t = t + l;
      if (fr) t = t + rt;
      if (fd) t = t + d;
      if (fu) t = t + u;
    }
    out.v[k] = t * pp.v[k];
  }
  __device__ __forceinline__ void end_row(size_t c, unsigned) { st4(q + c, out); }
};

__global__ void __launch_bounds__(TT) k_rb_forward(
    Grid g, TileList active, const double* __restrict__ r,
    const uint8_t* __restrict__ fluid, const double* __restrict__ precon, double* __restrict__ q,
    const DevScalars* sc) {
  if (sc->done) return;
  RbForward op{r, precon, q, {}, {}, {}, {}, {}};
  for_each_tile(g, active, [&](int x0, int y0, int y1, bool live) {
    stencil_tile(g, fluid, x0, y0, y1, live, op);
  });
}

// z = L^-T q.  Black: z = q*pc.  Red: z = (q + sum_nb pc*(q_nb*pc_nb)) * pc; window operand
// W = q*pc (z of a black neighbour).  Fused with the z.r reduction that follows every
// preconditioner application (main.c:748, 762) and the beta / sigma update (main.c:763-765).
struct RbBackward {
  const double* __restrict__ q;
  const double* __restrict__ pc;
  const double* __restrict__ r;
  double* __restrict__ z;
  double acc;
  int a0, a1;
  D4 qq, pp, qu, pu, rr, out;      // q, pc of the centre row / of the row above; r of the centre
  __device__ __forceinline__ D4 operand(size_t c, int slot) {
    const D4 a = ld4(q + c), b = ld4(pc + c);
    if (slot == SLOT_CENTER) { qq = a; pp = b; }
    if (slot == SLOT_UP) { qu = a; pu = b; }
    D4 w;
#pragma unroll
    for (int k = 0; k < 4; ++k) w.v[k] = a.v[k] * b.v[k];
    return w;
  }
  __device__ __forceinline__ double operand1(size_t c) const { return q[c] * pc[c]; }
  __device__ __forceinline__ void rotate() { qq = qu; pp = pu; }
  __device__ __forceinline__ void begin_row(size_t c, unsigned) {
    rr = ld4(r + c);
    out.v[0] = out.v[1] = out.v[2] = out.v[3] = 0.0;
  }
  __device__ __forceinline__ void cell(size_t, int k, double wc, double l, bool fl, double rt, bool fr,
                                       double d, bool fd, double u, bool fu, int x, int y) {
    double zc;
    if ((x + y) & 1) {
      zc = wc;                                               // black: q*pc
    } else {
      const double p = pp.v[k];
      double t = qq.v[k];
      if (fl) t = t + p * l;
      if (fr) t = t + p * rt;
      if (fd) t = t + p * d;
      if (fu) t = t + p * u;
      zc = t * p;
    }
    out.v[k] = zc;
    if (y >= a0 && y < a1) acc += zc * rr.v[k];
  }
  __device__ __forceinline__ void end_row(size_t c, unsigned) { st4(z + c, out); }   // single-GPU A/B variant only
};

__global__ void __launch_bounds__(TT) k_rb_backward(
    Grid g, TileList active, const double* __restrict__ q,
    const double* __restrict__ r, const uint8_t* __restrict__ fluid,
    const double* __restrict__ precon, double* __restrict__ z, double* partials, DevScalars* sc,
    int init, int exact, int acc0, int acc1) {
  if (sc->done) return;
  RbBackward op{q, precon, r, z, 0.0, acc0 + g.yoff, acc1 + g.yoff, {}, {}, {}, {}, {}, {}};
  for_each_tile(g, active, [&](int x0, int y0, int y1, bool live) {
    stencil_tile(g, fluid, x0, y0, y1, live, op);
  });
  const double bsum = block_reduce<false>(op.acc);
  grid_reduce_last_block<false>(bsum, partials, &sc->ctr[CTR_ZR], [&](double total) {
    if (exact == 2) { sc->part[0] = total; return; }
    if (exact) return;
    if (init) { sc->sigma = total; }                         // main.c:748
    else { sc->beta = total / sc->sigma; sc->sigma = total; }  // main.c:762-765
  });
}

// ---- reference-order dot product ---------------------------------------------------------
// dot() of main.c:629-639 sums a[y][x]*b[y][x] over fluid cells strictly in row-major order.
// fp64 addition is not associative, and the reference's solves are often unconverged at the
// 100-iteration cap (SURVEY §7 hard part 2), where 1-ulp differences in alpha/beta are
// amplified to ~1e-5 in p.  dot_mode = EULER_DOT_REFERENCE_ORDER reproduces the reference's
// sum bit for bit: the block forms the products of one 512-cell row segment at a time in
// shared memory (coalesced, double-buffered, skipping segments of tiles without fluid) and
// ONE thread chains the additions in order.  Adding +0.0 for a non-fluid cell never changes
// the running sum (it cannot be -0.0: it starts at +0.0).  Latency-bound by design:
// ~one dependent DADD per cell of the active tiles.
enum { DOT_ZS = 0, DOT_ZR_INIT = 1, DOT_ZR = 2 };

__global__ void __launch_bounds__(256) k_dot_seq(
    Grid g, const uint8_t* __restrict__ active, const double* __restrict__ a,
    const double* __restrict__ b, const uint8_t* __restrict__ fluid, DevScalars* sc, int what) {
  if (sc->done) return;
  __shared__ double buf[2][TW];
  const Tiles T = tiles_of(g);
  const long nseg = (long)g.ny * T.tx;
  auto next_active = [&](long k) {
    for (; k < nseg; ++k)
      if (active[(int)((k / T.tx) / g.th) * T.tx + (int)(k % T.tx)]) break;
    return k;
  };
  auto products = [&](long k, double& p0, double& p1) {
    const int y = (int)(k / T.tx), x = (int)(k % T.tx) * TW + threadIdx.x * 2;
    p0 = p1 = 0.0;
    if (x < g.pitch) {
      const size_t c = gidx(g, x, y);
      const double2 av = *reinterpret_cast<const double2*>(a + c);
      const double2 bv = *reinterpret_cast<const double2*>(b + c);
      const unsigned short m = *reinterpret_cast<const unsigned short*>(fluid + c);
      if (m & 0x00ffu) p0 = av.x * bv.x;
      if (m & 0xff00u) p1 = av.y * bv.y;
    }
  };
  double total = 0.0;
  long k = next_active(0);
  int cur = 0;
  if (k < nseg) {
    double p0, p1;
    products(k, p0, p1);
    buf[0][2 * threadIdx.x] = p0; buf[0][2 * threadIdx.x + 1] = p1;
  }
  __syncthreads();
  while (k < nseg) {
    const long kn = next_active(k + 1);
    double p0 = 0.0, p1 = 0.0;
    if (kn < nseg) products(kn, p0, p1);                     // loads in flight during the chain
    if (threadIdx.x == 0) {
      const double* v = buf[cur];
#pragma unroll 16
      for (int i = 0; i < TW; ++i) total += v[i];
    }
    buf[cur ^ 1][2 * threadIdx.x] = p0; buf[cur ^ 1][2 * threadIdx.x + 1] = p1;
    __syncthreads();
    cur ^= 1;
    k = kn;
  }
  if (threadIdx.x == 0) {
    if (what == DOT_ZS) { sc->zs = total; sc->alpha_prev = sc->alpha; sc->alpha = sc->sigma / total; }
    else if (what == DOT_ZR_INIT) { sc->sigma = total; }
    else { sc->beta = total / sc->sigma; sc->sigma = total; }
  }
}

// =========================================================================================
// TMA-fed variants of the three stencil kernels (pcg_pipe.cuh): same arithmetic, same
// results bit for bit, operands staged through a shared-memory ring by the bulk-copy engine.
// =========================================================================================
static_assert(pipe::TW == TW && pipe::TT == TT, "pipe tiling must match");

template <int NS>
__global__ void __launch_bounds__(TT) k_apply_a_pipe(
    Grid g, TileList active, const double* __restrict__ s,
    const uint8_t* __restrict__ fluid, const int8_t* __restrict__ adiag, double* __restrict__ z,
    double* partials, DevScalars* sc, int exact, int acc0, int acc1) {
  if (sc->done) return;
  ApplyAPipe op{g, z, 0.0, acc0, acc1};
  pipe::Planes<1, 2> in;
  in.d[0] = s; in.b[0] = fluid; in.b[1] = reinterpret_cast<const uint8_t*>(adiag);
  pipe::run<1, 2, NS, TH>(g, active.list, (int)*active.count, in, op);
  const double bsum = block_reduce<false>(op.acc);
  grid_reduce_last_block<false>(bsum, partials, &sc->ctr[CTR_ZS], [&](double total) {
    if (exact == 2) { sc->part[0] = total; return; }
    if (exact) return;
    sc->zs = total;
    sc->alpha_prev = sc->alpha; sc->alpha = sc->sigma / total;                           // main.c:752
  });
}

// MB: minimum resident blocks per SM asked of the compiler (0 = the tuned default of the fp64
// kernels); MB = 8 caps the fp32 instantiations at 64 registers (A/B knob EULER_MIXED_BLOCKS)
template <int NS, int C, class T = double, int MB = 0>
__global__ void __launch_bounds__(TW / C, MB ? MB : (C == 2 ? 4 : 5)) k_rb_forward_pipe(
    Grid g, TileList active, const T* __restrict__ r,
    const uint8_t* __restrict__ fluid, const T* __restrict__ precon, T* __restrict__ q,
    const DevScalars* sc) {
  pdl_prologue();
  if (sc->done) return;
  RbForwardPipe<C, T> op{g, q};
  pipe::Planes<2, 1, T> in;
  in.d[0] = r; in.d[1] = precon; in.b[0] = fluid;
  pipe::run<2, 1, NS, TH, RbForwardPipe<C, T>, C, T>(g, active.list, (int)*active.count, in, op);
}

template <int NS, int C, class T = double, int MB = 0>
__global__ void __launch_bounds__(TW / C, MB) k_rb_backward_pipe(
    Grid g, TileList active, const T* __restrict__ q,
    const T* __restrict__ r, const uint8_t* __restrict__ fluid,
    const T* __restrict__ precon, T* __restrict__ z, double* partials, DevScalars* sc,
    int init, int exact, int acc0, int acc1, double tol, const __grid_constant__ DistArgs dist) {
  pdl_prologue();
  if (sc->done) return;
  // (the peer planes are fp64: slabs run the fp64 solve; with T = float dist is all zeros)
  RbBackwardPipe<C, T> op{g, z, 0.0, acc0, acc1, reinterpret_cast<T*>(dist.z_dn), reinterpret_cast<T*>(dist.z_up),
                          dist.depth, false};
  pipe::Planes<3, 1, T> in;
  in.d[0] = q; in.d[1] = precon; in.d[2] = r; in.b[0] = fluid;
  pipe::run<3, 1, NS, TH, RbBackwardPipe<C, T>, C, T>(g, active.list, (int)*active.count, in, op);
  // peer stores of this thread are performed system-wide before the block reports in
  if (op.peer_stored) __threadfence_system();
  const double bsum = block_reduce<false>(op.acc);
  double total;
  if (!grid_reduce_last_block_all<false>(bsum, partials, &sc->ctr[CTR_ZR], total)) return;
  if (dist.mine) {                                           // halo flags + {z.r, ||r||inf} over NVLink
    p2p_finish(dist, sc, 1, init, tol, total, true);
    return;
  }
  if (threadIdx.x != 0) return;
  if (exact == 2) { sc->part[0] = total; return; }
  if (exact) return;
  if (init) { sc->sigma = total; }                           // main.c:748
  else { sc->beta = total / sc->sigma; sc->sigma = total; }  // main.c:762-765
}

// ---- slab mode: fold the all-gathered partials (rank order => deterministic) --------------
__global__ void k_dist_alpha(DevScalars* sc, const double* __restrict__ gathered, int nranks) {
  if (sc->done) return;
  double zs = 0.0;
  for (int r = 0; r < nranks; ++r) zs += gathered[r * 4 + 0];
  sc->zs = zs;
  sc->alpha_prev = sc->alpha;
  sc->alpha = sc->sigma / zs;                                // main.c:752
}
// after the preconditioner: z.r (sum) and, except for the initial application, ||r||inf (max)
// of the axpy that preceded it — the stop test of main.c:756 is evaluated here, one stage
// late: p and r are final either way, only the wasted z is different
__global__ void k_dist_beta(DevScalars* sc, const double* __restrict__ gathered, int nranks, int init,
                            double tol) {
  if (sc->done) return;
  double zr = 0.0, m = 0.0;
  for (int r = 0; r < nranks; ++r) { zr += gathered[r * 4 + 0]; m = fmax(m, gathered[r * 4 + 1]); }
  if (init) { sc->sigma = zr; return; }
  sc->resid = m;
  sc->iters += 1;
  if (m <= tol) { sc->done = 1; return; }
  sc->beta = zr / sc->sigma;
  sc->sigma = zr;
}

// ring depths: measured on B200 at 16384^2, shallower rings win (more resident blocks per SM
// hide the consumers' shared-memory latency better than a deeper prefetch does)
constexpr int NS_A = 4, NS_F = 4, NS_B = 5;

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// =========================================================================================
// Fused iteration for the red-black mode: TWO kernels per PCG iteration instead of five.
//
//   k_fused_search_apply   s' = z + beta s  (update_search, main.c:669-677)  and  A s'
//                          (apply_a, main.c:679-691) + the z.s reduction and alpha:
//                          34 B/cell instead of 24 + 18.
//   k_fused_axpy_forward   p += alpha s', r' = r - alpha A s' (main.c:753-754), ||r'||inf
//                          (:756) and q = L^-1 r' (first half of the red-black solve):
//                          65 B/cell instead of 48 + 25.  The second half (k_rb_backward,
//                          fused with z.r') follows; fusing it too would need r' on the
//                          radius-2 diamond of every cell (tried: the recomputation makes the
//                          kernel shared-memory/fp64 bound, profiles/r01 notes).
//
// Both recompute what a neighbour cell would have produced instead of reading it back from
// HBM: s' and r' on the 5-point halo of a cell.  Each recomputed value
// uses exactly the arithmetic of the unfused kernels, so results are bit-identical to them
// (and to the CPU mirror in the oracle).  Because neighbours are re-derived from the OLD s / r,
// the new s / r cannot be written in place: s and r ping-pong between two planes each.
// =========================================================================================

template <int NS, int C, class T = double, int MB = 0>
__global__ void __launch_bounds__(TW / C, MB) k_fused_search_apply(
    Grid g, TileList active, const T* __restrict__ z, const T* __restrict__ s,
    const uint8_t* __restrict__ fluid, const int8_t* __restrict__ adiag, T* __restrict__ s_new,
    T* __restrict__ as, double* partials, DevScalars* sc, int init, int exact, int acc0, int acc1,
    const __grid_constant__ DistArgs dist, int split_it, double tol, unsigned long long* tr) {
  pdl_prologue();
  if (sc->done) return;
  trace_mark(tr, 0);
  if (tr && blockIdx.x == 0 && threadIdx.x == 0) { tr[8] = 1; tr[9] = (unsigned long long)split_it; }
  double beta = sc->beta;
  if (split_it) {
    // split-phase exchange (p2p.cuh): this kernel's blocks consume {z.r, ||r||inf} that the
    // previous iteration's tail kernel posted — the stop test and beta of main.c:756-765
    if (split_it == 1) {
      if (blockIdx.x == 0 && threadIdx.x == 0) sc->sigma_s[0] = sc->sigma;   // z.r of the initial application
    } else {
      double zr, nm;
      const bool ok = p2p_collect(dist, true, zr, nm);
      const bool writer = blockIdx.x == 0 && threadIdx.x == 0;
      if (!ok) { if (writer) { sc->comm_timeout = 1; sc->done = 1; } return; }
      if (writer) { sc->resid = nm; sc->iters += 1; }
      if (nm <= tol) { if (writer) sc->done = 1; return; }                   // main.c:756-758
      beta = zr / sc->sigma_s[split_it & 1];                                // main.c:762-765
      if (writer) { sc->beta = beta; sc->sigma_s[(split_it + 1) & 1] = zr; }
    }
  }
  trace_mark(tr, 1);
  trace_block(tr, sc, 1, 0);
  FusedSearchApply<C, T> op{g, s_new, as, (T)beta, init != 0, 0.0, acc0, acc1};
  pipe::Planes<2, 2, T> in;
  in.d[0] = z; in.d[1] = s; in.b[0] = fluid; in.b[1] = reinterpret_cast<const uint8_t*>(adiag);
  pipe::run<2, 2, NS, TH, FusedSearchApply<C, T>, C, T>(g, active.list, (int)*active.count, in, op);
  trace_mark(tr, 2);
  trace_block(tr, sc, 1, 1);
  const double bsum = block_reduce<false>(op.acc);
  double total;
  if (!grid_reduce_last_block_all<false>(bsum, partials, &sc->ctr[CTR_ZS], total)) { trace_mark(tr, 3); return; }
  if (dist.mine) {                                           // {z.s} over NVLink -> alpha
    if (split_it) p2p_post(dist, total, 0.0, false);
    else p2p_finish(dist, sc, 0, 0, 0.0, total, false);
    trace_mark(tr, 3);
    return;
  }
  if (threadIdx.x != 0) return;
  if (exact == 2) { sc->part[0] = total; return; }
  sc->zs = total;
  sc->alpha_prev = sc->alpha; sc->alpha = sc->sigma / total;                             // main.c:752
  trace_mark(tr, 3);
}

template <int NS>
__global__ void __launch_bounds__(TT) k_true_residual(
    Grid g, TileList active, const double* __restrict__ p, const double* __restrict__ b,
    const uint8_t* __restrict__ fluid, const int8_t* __restrict__ adiag, float* __restrict__ r,
    const DevScalars* sc) {
  if (sc->done) return;
  TrueResidual op{g, b, r};
  pipe::Planes<1, 2> in;
  in.d[0] = p; in.b[0] = fluid; in.b[1] = reinterpret_cast<const uint8_t*>(adiag);
  pipe::run<1, 2, NS, TH>(g, active.list, (int)*active.count, in, op);
}

template <int NS, int C>
__global__ void __launch_bounds__(TW / C) k_fused_axpy_forward(
    Grid g, TileList active, const double* __restrict__ r, const double* __restrict__ as,
    const double* __restrict__ precon, const uint8_t* __restrict__ fluid,
    const double* __restrict__ s, double* __restrict__ p, double* __restrict__ r_new,
    double* __restrict__ q, double* partials, DevScalars* sc, double tol, int dist, int acc0,
    int acc1) {
  if (sc->done) return;
  FusedAxpyForward<C> op{g, s, p, r_new, q, sc->alpha, 0.0, acc0, acc1, {}, {}, ~(size_t)0};
  pipe::Planes<3, 1> in;
  in.d[0] = r; in.d[1] = as; in.d[2] = precon; in.b[0] = fluid;
  pipe::run<3, 1, NS, TH, FusedAxpyForward<C>, C>(g, active.list, (int)*active.count, in, op);
  const double bmax = block_reduce<true>(op.mx);
  grid_reduce_last_block<true>(bmax, partials, &sc->ctr[CTR_NORM], [&](double total) {
    if (dist) { sc->part[1] = total; return; }
    sc->resid = total;
    sc->iters += 1;
    if (total <= tol) sc->done = 1;                          // main.c:756-758
  });
}

#include "pcg_tail.cuh"

constexpr int NS_KA = 4, NS_KB = 4;
constexpr int CPT_F = 2, CPT_B = 2, CPT_KA = 4;   // cells per thread of the pipe kernels
// mixed-precision (fp32 storage) instantiations: 4 cells = one 16 B vector per thread and plane;
// ring depth per kernel in Ctx::ns_mixed (4, 6 or 8: a row segment is half the bytes of an fp64
// one, so the same bytes in flight need a deeper ring)
constexpr int CPT_MIXED = 4;

// persistent grid: resident blocks per SM (occupancy of that kernel) x SM count, capped by
// the number of tiles
template <class K>
int pcg_blocks(const Ctx& c, K kernel, int smem = 0, int threads = TT, bool max_carveout = false) {
  // occupancy is a property of (kernel, threads, smem) on this architecture: asked once per
  // kernel and host thread, not on every launch (two driver calls per launch were a visible
  // share of the host time per PCG iteration on thin slabs)
  // (the shared-memory opt-in is per device, hence the device in the key)
  thread_local std::map<std::pair<const void*, int>, int> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  const std::pair<const void*, int> key(reinterpret_cast<const void*>(kernel), dev);
  int per_sm = 0;
  auto it = cache.find(key);
  if (it != cache.end()) {
    per_sm = it->second;
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (max_carveout)             // 8 blocks x ~28 KB need (nearly) all of the SM's shared memory
      cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    cache[key] = per_sm;
  }
  // at most one block per 8 rows of a tile column, so that tiny grids do not pay two halo
  // rows per output row
  const Tiles T = tiles_of(c.g);
  // A/B knob for thin slabs: EULER_PCG_MAX_BLOCKS_PER_SM caps the resident
  // blocks of every persistent PCG kernel, i.e. fewer, longer row ranges per block
  static const int max_per_sm = env_int("EULER_PCG_MAX_BLOCKS_PER_SM", 0);
  if (max_per_sm > 0 && per_sm > max_per_sm) per_sm = max_per_sm;
  const long want = (long)c.sm_count * per_sm, cap = (long)T.n * c.g.th / 8;
  return (int)(cap < 1 ? 1 : (cap < want ? cap : want));
}

}  // namespace

// The PCG kernels work on the rows this handle OWNS: a view with advanced base pointers, so
// that row -1 and row ny of the view are the halo rows received from the neighbouring slabs
// (or guard rows of zeros on a single GPU).
struct PV {
  Grid g;
  int a0, a1;            // the OWNED rows, in view coordinates: only they enter reductions
  const uint8_t* fluid;
  const int8_t* adiag;
  double *s, *z, *r, *p, *q, *precon;
  float *s32, *z32, *r32, *q32, *pc32;     // mixed-precision mode only (null otherwise)
};
// Slab mode: the view is the owned rows +-PCG_EXT halo rows.  Every vector update is done
// redundantly on those halo rows with the owner's exact arithmetic, so one exchange of s
// (4 rows) per iteration keeps r, p, q, z consistent without exchanging them: validity
// shrinks by one row per stencil (s: +-4 -> A s, r: +-3 -> q: +-2 -> z, new s: +-1).
constexpr int PCG_EXT = 3;
static PV pview(const Ctx& c) {
  const int ext = c.distributed ? PCG_EXT : 0;
  const int lo = c.own0 - ext > 0 ? c.own0 - ext : 0;
  const int hi = c.own1 + ext < c.g.ny ? c.own1 + ext : c.g.ny;
  const size_t o = (size_t)lo * c.g.pitch;
  PV v;
  v.g = c.g; v.g.ny = hi - lo; v.g.yoff = c.g.yoff + lo;
  v.a0 = c.own0 - lo; v.a1 = c.own1 - lo;
  // tiles are only the unit of the "contains fluid" flags; the persistent kernels split the
  // active tiles' ROWS evenly over their blocks, so a thin slab needs no smaller tile
  // (EULER_PCG_TH: A/B knob — taller tiles halve the halo rows the fused tail recomputes per tile)
  static const int pcg_th = env_int("EULER_PCG_TH", 2 * TH);
  v.g.th = pcg_th;
  v.fluid = c.count + o; v.adiag = c.adiag + o;
  v.s = c.s + o; v.z = c.z + o; v.r = c.r + o; v.p = c.p + o; v.q = c.q + o; v.precon = c.precon + o;
  v.s32 = v.z32 = v.r32 = v.q32 = v.pc32 = nullptr;
  if (c.mixed) { v.s32 = c.s32 + o; v.z32 = c.z32 + o; v.r32 = c.r32 + o; v.q32 = c.q32 + o; v.pc32 = c.pc32 + o; }
  return v;
}
static inline int dotflag(const Ctx& c) { return c.distributed ? 2 : c.dot_mode; }

void launch_dot_zr_exact(Ctx& c, bool init);

void launch_dist_alpha(Ctx& c, const double* gathered, int nranks) {
  k_dist_alpha<<<1, 1, 0, c.stream>>>(c.sc, gathered, nranks);
  c.launches += 1;
}
void launch_dist_beta(Ctx& c, const double* gathered, int nranks, bool init, double tol) {
  k_dist_beta<<<1, 1, 0, c.stream>>>(c.sc, gathered, nranks, init ? 1 : 0, tol);
  c.launches += 1;
}

void launch_tile_flags(Ctx& c) {
  ProfScope ps(c, KC_MISC);
  const PV v = pview(c);
  const Tiles T = tiles_of(v.g);
  const int nb = T.n < c.sm_count * 16 ? T.n : c.sm_count * 16;
  k_tile_flags<<<nb, TT, 0, c.stream>>>(v.g, v.fluid, c.tile_active, c.sc);
  k_tile_compact<<<1, 1024, 0, c.stream>>>(v.g, c.tile_active, c.tile_list, c.sc);
  c.launches += 2;
}

void launch_pcg_reset(Ctx& c) {
  ProfScope ps(c, KC_MISC);
  k_pcg_reset<<<1, 1, 0, c.stream>>>(c.sc);
  c.launches += 1;
}

#define TL TileList{c.tile_list, &c.sc->active_tiles}

void launch_apply_a(Ctx& c, bool) {
  ProfScope ps(c, KC_APPLY_A);
  const PV v = pview(c);
  if (c.use_pipe) {
    constexpr int smem = pipe::smem_bytes<1, 2, NS_A>();
    k_apply_a_pipe<NS_A><<<pcg_blocks(c, k_apply_a_pipe<NS_A>, smem), TT, smem, c.stream>>>(
        v.g, TL, v.s, v.fluid, v.adiag, v.z, c.partials, c.sc, dotflag(c), v.a0, v.a1);
  } else {
    k_apply_a<<<pcg_blocks(c, k_apply_a), TT, 0, c.stream>>>(v.g, TL, v.s, v.fluid, v.adiag, v.z,
                                                             c.partials, c.sc, dotflag(c), v.a0, v.a1);
  }
  c.launches += 1;
  if (c.dot_mode && !c.distributed) {
    k_dot_seq<<<1, 256, 0, c.stream>>>(v.g, c.tile_active, v.z, v.s, v.fluid, c.sc, DOT_ZS);
    c.launches += 1;
  }
}

void launch_axpy(Ctx& c, double tol, bool as_in_q, int mode) {
  ProfScope ps(c, KC_AXPY);
  const PV v = pview(c);
  const size_t o = (size_t)(v.s - c.s);
  if (c.mixed)                    // fused red-black iteration only: A s is in q32
    launch_pdl(k_axpy<float>, pcg_blocks(c, k_axpy<float>), TT, 0, c.stream, v.g, TL, v.s32, c.s32b + o,
               v.q32, v.fluid, v.p, v.r32, c.partials, c.sc, tol, c.distributed ? 1 : 0, v.a0, v.a1, mode);
  else
    launch_pdl(k_axpy<double>, pcg_blocks(c, k_axpy<double>), TT, 0, c.stream, v.g, TL, v.s, c.s2 ? c.s2 + o : v.s,
               as_in_q ? v.q : v.z, v.fluid, v.p, v.r, c.partials, c.sc, tol, c.distributed ? 1 : 0, v.a0, v.a1,
               mode);
  c.launches += 1;
}

void launch_p_fixup(Ctx& c, const void* s_odd_plane, int split) {
  ProfScope ps(c, KC_MISC);
  const PV v = pview(c);
  const size_t o = (size_t)(v.s - c.s);
  if (c.mixed)
    k_p_fixup<float><<<pcg_blocks(c, k_p_fixup<float>), TT, 0, c.stream>>>(
        v.g, TL, static_cast<const float*>(s_odd_plane) + o, v.fluid, v.p, c.sc, split);
  else
    k_p_fixup<double><<<pcg_blocks(c, k_p_fixup<double>), TT, 0, c.stream>>>(
        v.g, TL, static_cast<const double*>(s_odd_plane) + o, v.fluid, v.p, c.sc, split);
  c.launches += 1;
}

// mixed-precision mode: r32 <- b - A p in fp64 (b is what k_build_rhs left in the fp64 r plane,
// which this mode never updates)
void launch_true_residual(Ctx& c) {
  ProfScope ps(c, KC_RESIDUAL);
  const PV v = pview(c);
  constexpr int smem = pipe::smem_bytes<1, 2, NS_A>();
  k_true_residual<NS_A><<<pcg_blocks(c, k_true_residual<NS_A>, smem), TT, smem, c.stream>>>(
      v.g, TL, v.p, v.r, v.fluid, v.adiag, v.r32, c.sc);
  c.launches += 1;
}

void launch_update_search(Ctx& c) {
  ProfScope ps(c, KC_UPDATE_SEARCH);
  const PV v = pview(c);
  k_update_search<<<pcg_blocks(c, k_update_search), TT, 0, c.stream>>>(v.g, TL, v.z, v.fluid, v.s, c.sc);
  c.launches += 1;
}

void launch_copy_search(Ctx& c) {
  ProfScope ps(c, KC_MISC);
  const PV v = pview(c);
  k_copy_search<<<pcg_blocks(c, k_copy_search), TT, 0, c.stream>>>(v.g, TL, v.z, v.s);
  c.launches += 1;
}

void launch_rb_build(Ctx& c) {
  // over ALL locally stored rows: halo rows get their own (identical) factor, no exchange
  ProfScope ps(c, KC_PRECON_BUILD);
  const dim3 block(32, 8);
  const long pieces = (long)c.gt_tx * c.gt_ty * GT_SUB, want = (long)c.sm_count * 8;
  const int grid = (int)(pieces < want ? pieces : want);
  const GridTiles gt = grid_tiles_of(c, c.gt_sparse);
  if (c.mixed) k_rb_build<float><<<grid, block, 0, c.stream>>>(c.g, gt, c.count, c.adiag, c.pc32);
  else k_rb_build<double><<<grid, block, 0, c.stream>>>(c.g, gt, c.count, c.adiag, c.precon);
  c.launches += 1;
}

void launch_rb_forward(Ctx& c) {
  ProfScope ps(c, KC_PRECON_FWD);
  const PV v = pview(c);
  if (c.mixed) {
    constexpr int C = CPT_MIXED;
#define FWD32(N) { constexpr int sf = pipe::smem_bytes<2, 1, N, float>(); \
    launch_pdl(k_rb_forward_pipe<N, C, float>, pcg_blocks(c, k_rb_forward_pipe<N, C, float>, sf, TW / C), \
               TW / C, sf, c.stream, v.g, TL, v.r32, v.fluid, v.pc32, v.q32, c.sc); }
    if (c.mixed_blocks == 8) {      // 64-register instantiation, 8 resident blocks (ring depth 4)
      constexpr int sf = pipe::smem_bytes<2, 1, 4, float>();
      auto k = k_rb_forward_pipe<4, C, float, 8>;
      launch_pdl(k, pcg_blocks(c, k, sf, TW / C, true), TW / C, sf, c.stream, v.g, TL, v.r32, v.fluid, v.pc32, v.q32, c.sc);
    } else if (c.ns_mixed[0] == 8) FWD32(8) else if (c.ns_mixed[0] == 6) FWD32(6) else FWD32(4)
#undef FWD32
  } else if (c.use_pipe) {
    static const int ns = env_int("EULER_NS_F", NS_F);
    static const int cpt = env_int("EULER_CPT_F", CPT_F);
#define FWD(N, C) { constexpr int sf = pipe::smem_bytes<2, 1, N>(); \
    launch_pdl(k_rb_forward_pipe<N, C>, pcg_blocks(c, k_rb_forward_pipe<N, C>, sf, TW / C), TW / C, sf, c.stream, \
               v.g, TL, v.r, v.fluid, v.precon, v.q, c.sc); }
    if (cpt == 2) { if (ns == 6) FWD(6, 2) else if (ns == 5) FWD(5, 2) else FWD(4, 2) }
    else { if (ns == 6) FWD(6, 4) else if (ns == 5) FWD(5, 4) else FWD(4, 4) }
#undef FWD
  } else {
    k_rb_forward<<<pcg_blocks(c, k_rb_forward), TT, 0, c.stream>>>(v.g, TL, v.r, v.fluid, v.precon, v.q, c.sc);
  }
  c.launches += 1;
}

void launch_rb_backward(Ctx& c, bool init) {
  ProfScope ps(c, KC_PRECON_BWD);
  const PV v = pview(c);
  if (c.mixed) {
    constexpr int C = CPT_MIXED;
    DistArgs d;
    memset(&d, 0, sizeof d);
    if (c.p2p_mode == 2) {        // NVLink path on slabs: the neighbours' fp32 z planes, biased in BYTES
      d = c.dist;
      d.depth = P2P_HALO_DEPTH;
      const long lo = (long)((v.z32 - c.z32) / c.g.pitch);
      if (d.z_dn) d.z_dn = reinterpret_cast<double*>(reinterpret_cast<char*>(d.z_dn) + (lo + c.p2p_dn_own1 - c.own0) * (long)c.g.pitch * 4);
      if (d.z_up) d.z_up = reinterpret_cast<double*>(reinterpret_cast<char*>(d.z_up) + (lo + c.p2p_up_own0 - c.own1) * (long)c.g.pitch * 4);
    }
#define BWD32(N) { constexpr int sb = pipe::smem_bytes<3, 1, N, float>(); \
    launch_pdl(k_rb_backward_pipe<N, C, float>, pcg_blocks(c, k_rb_backward_pipe<N, C, float>, sb, TW / C), \
               TW / C, sb, c.stream, v.g, TL, v.q32, v.r32, v.fluid, v.pc32, v.z32, c.partials, c.sc, init ? 1 : 0, \
               dotflag(c), v.a0, v.a1, c.tol, d); }
    if (c.mixed_blocks == 8) {
      constexpr int sb = pipe::smem_bytes<3, 1, 4, float>();
      auto k = k_rb_backward_pipe<4, C, float, 8>;
      launch_pdl(k, pcg_blocks(c, k, sb, TW / C, true), TW / C, sb, c.stream, v.g, TL, v.q32, v.r32, v.fluid, v.pc32,
                 v.z32, c.partials, c.sc, init ? 1 : 0, dotflag(c), v.a0, v.a1, c.tol, d);
    } else if (c.ns_mixed[1] == 8) BWD32(8) else if (c.ns_mixed[1] == 6) BWD32(6) else BWD32(4)
#undef BWD32
  } else if (c.use_pipe) {
    static const int ns = env_int("EULER_NS_B", NS_B);
    static const int cpt = env_int("EULER_CPT_B", CPT_B);
    // NVLink path: this kernel also stores its edge rows into the neighbours' halo rows and its
    // last block finishes {z.r, ||r||inf} across ranks (p2p.cuh).  The neighbours' planes are
    // biased so that gidx(view, x, y) of an owned edge row addresses the matching halo row:
    // view row y = local row y + lo; my first owned row <-> their first row above their owned
    // rows, my last owned row <-> their last row below their owned rows.
    DistArgs d;
    memset(&d, 0, sizeof d);
    if (c.p2p_mode == 2) {
      d = c.dist;
      d.depth = P2P_HALO_DEPTH;
      const long lo = (long)((v.z - c.z) / c.g.pitch);
      if (d.z_dn) d.z_dn += (lo + c.p2p_dn_own1 - c.own0) * (long)c.g.pitch;
      if (d.z_up) d.z_up += (lo + c.p2p_up_own0 - c.own1) * (long)c.g.pitch;
    }
#define BWD(N, C) { constexpr int sb = pipe::smem_bytes<3, 1, N>(); \
    launch_pdl(k_rb_backward_pipe<N, C>, pcg_blocks(c, k_rb_backward_pipe<N, C>, sb, TW / C), TW / C, sb, c.stream, \
               v.g, TL, v.q, v.r, v.fluid, v.precon, v.z, c.partials, c.sc, init ? 1 : 0, dotflag(c), v.a0, v.a1, \
               c.tol, d); }
    if (cpt == 2) { if (ns == 4) BWD(4, 2) else if (ns == 6) BWD(6, 2) else BWD(5, 2) }
    else { if (ns == 4) BWD(4, 4) else if (ns == 6) BWD(6, 4) else BWD(5, 4) }
#undef BWD
  } else {
    k_rb_backward<<<pcg_blocks(c, k_rb_backward), TT, 0, c.stream>>>(
        v.g, TL, v.q, v.r, v.fluid, v.precon, v.z, c.partials, c.sc, init ? 1 : 0, dotflag(c), v.a0, v.a1);
  }
  c.launches += 1;
  launch_dot_zr_exact(c, init);
}

void launch_rb_apply(Ctx& c, bool init) {
  launch_rb_forward(c);
  launch_rb_backward(c, init);
}

// fused iteration (see k_fused_*).  Invariant at the start of an iteration: c.z = M^-1 r.
//   search_apply : reads c.z, s            writes s' (twin plane, swapped in), A s' -> c.q
//   axpy_forward : reads A s' (c.q), r, pc writes p, r' (twin, swapped in), q -> c.z; then
//                  c.z <-> c.q so that the ordinary k_rb_backward reads q from c.q and leaves
//                  the new M^-1 r' in c.z
void launch_fused_search_apply(Ctx& c, bool init, int split_it) {
  ProfScope ps(c, KC_FUSED_A);
  const PV v = pview(c);
  const size_t o = (size_t)(v.s - c.s);
  static const int ns = env_int("EULER_NS_KA", NS_KA);
  static const int cpt = env_int("EULER_CPT_KA", CPT_KA);
  DistArgs d;                     // NVLink path: the last block finishes {z.s} across ranks
  memset(&d, 0, sizeof d);
  if (c.mixed) {
    constexpr int C = CPT_MIXED;
    if (c.p2p_mode == 2) d = c.dist;   // slabs, NVLink path: {z.s} finished across ranks in the last block
    const int xf = c.distributed ? 2 : 0;
#define KA32(N) { constexpr int smem = pipe::smem_bytes<2, 2, N, float>(); \
    launch_pdl(k_fused_search_apply<N, C, float>, \
               pcg_blocks(c, k_fused_search_apply<N, C, float>, smem, TW / C), TW / C, smem, c.stream, \
               v.g, TL, v.z32, v.s32, v.fluid, v.adiag, c.s32b + o, v.q32, c.partials, c.sc, init ? 1 : 0, xf, \
               v.a0, v.a1, d, 0, 0.0, (unsigned long long*)nullptr); }
    if (c.mixed_blocks == 8) {
      constexpr int smem = pipe::smem_bytes<2, 2, 4, float>();
      auto k = k_fused_search_apply<4, C, float, 8>;
      launch_pdl(k, pcg_blocks(c, k, smem, TW / C, true), TW / C, smem, c.stream, v.g, TL, v.z32, v.s32, v.fluid,
                 v.adiag, c.s32b + o, v.q32, c.partials, c.sc, init ? 1 : 0, xf, v.a0, v.a1, d, 0, 0.0, (unsigned long long*)nullptr);
    } else if (c.ns_mixed[2] == 8) KA32(8) else if (c.ns_mixed[2] == 6) KA32(6) else KA32(4)
#undef KA32
    c.launches += 1;
    float* t32 = c.s32; c.s32 = c.s32b; c.s32b = t32;
    return;
  }
  if (c.p2p_mode == 2) d = c.dist;
  unsigned long long* tr = trace_slot(c);
#define KA(N, C) { constexpr int smem = pipe::smem_bytes<2, 2, N>(); \
  launch_pdl(k_fused_search_apply<N, C>, pcg_blocks(c, k_fused_search_apply<N, C>, smem, TW / C), TW / C, smem, c.stream, \
             v.g, TL, v.z, v.s, v.fluid, v.adiag, c.s2 + o, v.q, c.partials, c.sc, init ? 1 : 0, \
             c.distributed ? 2 : 0, v.a0, v.a1, d, split_it, c.tol, tr); }
  if (cpt == 2) { if (ns == 10) KA(10, 2) else if (ns == 8) KA(8, 2) else if (ns == 6) KA(6, 2) else if (ns == 5) KA(5, 2) else KA(4, 2) }
  else { if (ns == 6) KA(6, 4) else if (ns == 5) KA(5, 4) else KA(4, 4) }
#undef KA
  c.launches += 1;
  double* t = c.s; c.s = c.s2; c.s2 = t;
}

void launch_fused_axpy_forward(Ctx& c, double tol) {
  ProfScope ps(c, KC_FUSED_B);
  const PV v = pview(c);
  const size_t o = (size_t)(v.r - c.r);
  constexpr int smem = pipe::smem_bytes<3, 1, NS_KB>();
  constexpr int C = 2;
  k_fused_axpy_forward<NS_KB, C><<<pcg_blocks(c, k_fused_axpy_forward<NS_KB, C>, smem, TW / C), TW / C, smem, c.stream>>>(
      v.g, TL, v.r, v.q, v.precon, v.fluid, v.s, v.p, c.r2 + o, v.z, c.partials, c.sc, tol,
      c.distributed ? 1 : 0, v.a0, v.a1);
  c.launches += 1;
  double* t = c.r; c.r = c.r2; c.r2 = t;
  t = c.z; c.z = c.q; c.q = t;
}

// axpy + forward + backward of the fused red-black iteration as one kernel (pcg_tail.cuh):
// reads r, A s (c.q), pc, s (, s_prev), p; writes r' into the twin plane (swapped in), p, z
void launch_fused_tail(Ctx& c, double tol, int mode, int split_it) {
  ProfScope ps(c, KC_FUSED_TAIL);
  const PV v = pview(c);
  const size_t o = (size_t)(v.r - c.r);
  DistArgs d;
  memset(&d, 0, sizeof d);
  if (c.p2p_mode == 2) {          // NVLink path: edge rows of z go straight into the neighbours' halo rows
    d = c.dist;
    d.depth = P2P_HALO_DEPTH;
    const long lo = (long)((v.z - c.z) / c.g.pitch);
    if (d.z_dn) d.z_dn += (lo + c.p2p_dn_own1 - c.own0) * (long)c.g.pitch;
    if (d.z_up) d.z_up += (lo + c.p2p_up_own0 - c.own1) * (long)c.g.pitch;
  }
  unsigned long long* tr = trace_slot(c);
  static const int cpt = env_int("EULER_CPT_TAIL", 2);
  static const int ns = env_int("EULER_NS_TAIL", 3);
  // resident blocks per SM the compiler must allow.  Measured at 16384^2 (same box, profiles/r02a):
  // 3 blocks = 72 registers with 48 B of spills inside the row loop, 1.12 ms per launch; 2 blocks =
  // 96 registers, no spill, 0.636 ms (ring depth 3 / 4 / 5: 0.636 / 0.639 / 0.648 ms)
  static const int mb = env_int("EULER_TAIL_MINB", 2);
#define TAIL(N, C, MB) { constexpr int smem = tail::smem_bytes<N>(); constexpr int threads = TW / C + 32; \
    k_fused_tail<N, C, MB><<<pcg_blocks(c, k_fused_tail<N, C, MB>, smem, threads), threads, smem, c.stream>>>( \
        v.g, TL, v.r, v.q, v.precon, v.fluid, v.s, c.s2 + o, v.p, c.r2 + o, v.z, c.partials, c.sc, tol, mode, \
        dotflag(c), v.a0, v.a1, d, split_it, tr); }
  if (cpt == 4) { if (ns == 4) TAIL(4, 4, 3) else TAIL(3, 4, 3) }
  else if (mb == 2) { if (ns == 5) TAIL(5, 2, 2) else if (ns == 4) TAIL(4, 2, 2) else TAIL(3, 2, 2) }
  else { if (ns == 4) TAIL(4, 2, 3) else TAIL(3, 2, 3) }
#undef TAIL
  c.launches += 1;
  double* t = c.r; c.r = c.r2; c.r2 = t;
}

__global__ void k_set_alpha(DevScalars* sc, double alpha) {
  sc->alpha = alpha; sc->alpha_prev = 0.0; sc->sigma = 1.0;
}
void launch_set_alpha(Ctx& c, double alpha) {
  k_set_alpha<<<1, 1, 0, c.stream>>>(c.sc, alpha);
  c.launches += 1;
}

// Split-phase exchange: what the LAST tail kernel of a batch posted has no consumer yet.  This
// one-block kernel reads it for the host (peek_*: stop test and iteration count as they will
// be) and, when the host ends the solve (`apply`), makes it the solve's final state.
__global__ void __launch_bounds__(64) k_dist_peek(DevScalars* sc, double tol, int apply,
                                                  const __grid_constant__ DistArgs dist) {
  if (sc->done) { if (threadIdx.x == 0) { sc->peek_done = 1; sc->peek_iters = sc->iters; sc->peek_resid = sc->resid; } return; }
  double zr, nm;
  const bool ok = p2p_collect(dist, true, zr, nm);
  if (threadIdx.x != 0) return;
  if (!ok) { sc->comm_timeout = 1; sc->done = 1; return; }
  sc->peek_resid = nm; sc->peek_iters = sc->iters + 1; sc->peek_done = nm <= tol ? 1 : 0;
  if (apply) { sc->resid = nm; sc->iters += 1; if (nm <= tol) sc->done = 1; }
}
void launch_dist_peek(Ctx& c, bool apply) {
  k_dist_peek<<<1, 64, 0, c.stream>>>(c.sc, c.tol, apply ? 1 : 0, c.dist);
  c.launches += 1;
}

void launch_dot_zr_exact(Ctx& c, bool init) {
  if (!c.dot_mode || c.distributed) return;
  const PV v = pview(c);
  k_dot_seq<<<1, 256, 0, c.stream>>>(v.g, c.tile_active, v.z, v.r, v.fluid, c.sc,
                                     init ? DOT_ZR_INIT : DOT_ZR);
  c.launches += 1;
}

// capacity for the smallest tile height
int pcg_tile_count(const Grid& g) { Grid t = g; t.th = 4; return tiles_of(t).n; }
int pcg_tile_cells(const Ctx& c) { return TW * pview(c).g.th; }

}  // namespace euler
