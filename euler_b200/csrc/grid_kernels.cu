// euler_b200/csrc/grid_kernels.cu — the MAC-grid (non-PCG) stages of the sub-step.
//
//   k_maxsq + k_timestep      calculate_timestep / maxsq          reference main.c:808-841
//   k_extrapolate_bounds      extrapolate(u),(v) + zero_bounds    main.c:158-185, 822-832, 865-868
//   k_advect_velocity         advect_u, advect_v, apply_body_forces, zero_bounds(tmp)
//                                                                 main.c:382-422, 539-545, 871-889
//   k_build_rhs               b, a_diag, p=0 (first part of project) main.c:713-733, 739
//   k_pressure_update         clamp p, subtract grad p            main.c:769-805
//
// All of them are HBM-bound streaming stencils (DESIGN.md §4 has the byte counts).  A thread
// owns FOUR consecutive P cells of a row, with the U faces on their right and the V faces
// above them: masks come in as one 32-bit word per plane and row, fp32 planes as float4, fp64
// planes as 2 x double2, so a warp moves 128 B / 512 B / 1 KiB per instruction, and a quad
// whose neighbourhood holds no fluid (most of a free-surface scene) costs two mask loads and
// two 16 B stores.  The one-cell-per-thread forms (`*_scalar`, EULER_GRID_VARIANT=1) are kept
// for A/B runs: they were issue-bound on byte loads at 33-42 % of the HBM peak (profiles/r01c).
#include <stdlib.h>

#include "interp.cuh"
#include "kernels.h"
#include "grid_ops.cuh"

namespace euler {

namespace {

constexpr int BX = 32, BY = 8;

inline dim3 grid2d(const Grid& g) { return dim3((g.nx + BX - 1) / BX, (g.ny + BY - 1) / BY); }

// ------------------------------------------------------------------ timestep ----

// max over ALL U faces of u^2 and ALL V faces of v^2 (main.c:808-820: air/solid included).
// `value > max` is false for NaN, so NaNs are skipped exactly like the reference does.
__global__ void __launch_bounds__(256) k_maxsq(Grid g, int r0, int r1, const float* __restrict__ u,
                                                const float* __restrict__ v, DevScalars* sc) {
  float mu = 0.f, mv = 0.f;
  const int quads = g.pitch >> 2;
  const size_t total = (size_t)quads * (r1 - r0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int y = r0 + (int)(i / quads);
    const int x0 = (int)(i % quads) << 2;
    const float4 a = *reinterpret_cast<const float4*>(u + gidx(g, x0, y));
    const float4 b = *reinterpret_cast<const float4*>(v + gidx(g, x0, y));
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + k;
      if (x < g.nx - 1) { float s = av[k] * av[k]; if (s > mu) mu = s; }
      if (x < g.nx && y + g.yoff < g.gny - 1) { float s = bv[k] * bv[k]; if (s > mv) mv = s; }
    }
  }
  mu = warp_maxf(mu);
  mv = warp_maxf(mv);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&sc->max_u2_bits, __float_as_uint(mu));
    atomicMax(&sc->max_v2_bits, __float_as_uint(mv));
  }
}

// dt = fminf(0.75h / sqrtf(max u^2 + max v^2), frame_time)   (main.c:838-840)
__global__ void k_timestep(DevScalars* sc, float reach, float frame_time) {
  const float vmax = sqrtf(__uint_as_float(sc->max_u2_bits) + __uint_as_float(sc->max_v2_bits));
  sc->dt = fminf(reach / vmax, frame_time);
}

// --------------------------------------------------------- extrapolate+bounds ----

template <int TYPE>
__device__ __forceinline__ float extrapolated_face(const Grid& g, const float* __restrict__ q,
                                                   const uint8_t* __restrict__ fluid,
                                                   const uint8_t* __restrict__ prev,
                                                   const uint8_t* __restrict__ solid, int x, int y) {
  // sizes and the clamped 3x3 block are GLOBAL notions (main.c:179-180); y is a view row
  const int sx = g.nx - (TYPE == FACE_U), sy = g.gny - (TYPE == FACE_V) - g.yoff;
  const bool now = face_has<TYPE>(fluid, g, x, y);
  // zero_bounds (main.c:827): not touching fluid, or touching a solid -> 0
  if (!now || face_has<TYPE>(solid, g, x, y)) return 0.f;
  float val = q[gidx(g, x, y)];
  if (!face_has<TYPE>(prev, g, x, y)) {
    // newly wet face: mean of the clamped 3x3 block's faces that were wet (main.c:158-171,
    // 179-181); row-major accumulation order; 0/0 -> NaN when there is none (assert is off)
    const int x0 = max(x - 1, 0), x1 = min(x + 1, sx - 1);
    const int y0 = max(y - 1, -g.yoff), y1 = min(y + 1, sy - 1);
    float total = 0.f;
    int n = 0;
    for (int yy = y0; yy <= y1; ++yy)
      for (int xx = x0; xx <= x1; ++xx)
        if (face_has<TYPE>(prev, g, xx, yy)) { total += q[gidx(g, xx, yy)]; ++n; }
    val = total / (float)n;
  }
  return val;
}

// Out of place: a face that stops being wet is zeroed here while a neighbour may still need
// its old value for the 3x3 mean (in the reference the two passes are sequential).
__global__ void __launch_bounds__(BX* BY) k_extrapolate_bounds_scalar(
    Grid g, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ prev,
    const uint8_t* __restrict__ solid, float* __restrict__ uo, float* __restrict__ vo) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  if (x >= g.nx || y >= g.ny) return;
  const size_t c = gidx(g, x, y);
  uo[c] = x < g.nx - 1 ? extrapolated_face<FACE_U>(g, u, fluid, prev, solid, x, y) : 0.f;
  vo[c] = y + g.yoff < g.gny - 1 ? extrapolated_face<FACE_V>(g, v, fluid, prev, solid, x, y) : 0.f;
}

// ------------------------------------------------------------ velocity advect ----

__global__ void __launch_bounds__(BX* BY) k_advect_velocity_scalar(
    Grid g, InterpLimits lim, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid,
    float* __restrict__ uo, float* __restrict__ vo, float dt, float h, float gravity) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  if (x >= g.nx || y >= g.ny) return;
  const size_t c = gidx(g, x, y);
  const int gy = y + g.yoff;                 // sample positions are in global index space
  float ru = 0.f, rv = 0.f;
  if (x < g.nx - 1 && face_has<FACE_U>(fluid, g, x, y) && !face_has<FACE_U>(solid, g, x, y)) {
    // main.c:388-395: back-trace one Euler step, sample u there
    const float dx = u[c];
    const float dy = interpolate<FACE_V>(v, fluid, g, lim, x + 0.5f, gy - 0.5f);
    const float px = x - div_h(dx * dt, h);
    const float py = gy - div_h(dy * dt, h);
    ru = interpolate<FACE_U>(u, fluid, g, lim, px, py);
  }
  if (gy < g.gny - 1 && face_has<FACE_V>(fluid, g, x, y) && !face_has<FACE_V>(solid, g, x, y)) {
    // main.c:411-418, then gravity main.c:542
    const float dy = v[c];
    const float dx = interpolate<FACE_U>(u, fluid, g, lim, x - 0.5f, gy + 0.5f);
    const float px = x - div_h(dx * dt, h);
    const float py = gy - div_h(dy * dt, h);
    rv = interpolate<FACE_V>(v, fluid, g, lim, px, py);
    rv += gravity * dt;
  }
  uo[c] = ru;
  vo[c] = rv;
}

// ------------------------------------------------------------------ rhs build ----

__global__ void __launch_bounds__(BX* BY) k_build_rhs_scalar(
    Grid g, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid,
    double* __restrict__ r, double* __restrict__ p, int8_t* __restrict__ adiag, float h,
    double scale, DevScalars* sc, int own0, int own1, float* __restrict__ r32) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  bool nz = false;
  if (x < g.nx && y < g.ny) {
    const size_t c = gidx(g, x, y);
    double b = 0.0;
    if (fluid[c]) {
      // main.c:720-721: divergence left to right in fp32, widened, scaled by h^2 rho/dt
      const float div = div_h(u[c] - u[c - 1] + v[c] - v[c - g.pitch], h);
      b = -(double)div * scale;
      // main.c:554-559: 4 minus the number of solid neighbours
      adiag[c] = (int8_t)(4 - solid[c - 1] - solid[c + 1] - solid[c - g.pitch] - solid[c + g.pitch]);
      nz = (b != 0.0) && y >= own0 && y < own1;   // halo rows are the neighbour slab's business
    }
    r[c] = b;
    p[c] = 0.0;
    if (r32) r32[c] = (float)b;                              // mixed-precision PCG: r starts as fp32(b)
  }
  if (__any_sync(EULER_FULL_MASK, nz) && (threadIdx.x & 31) == 0) atomicOr(&sc->nonzero_rhs, 1);
}

// ------------------------------------------------------------ pressure update ----

__device__ __forceinline__ double clamped_p(const double* __restrict__ p,
                                            const uint8_t* __restrict__ fluid, size_t c) {
  double v = p[c];
  return (fluid[c] && v < 0.0) ? 0.0 : v;                    // main.c:773-779
}

__global__ void __launch_bounds__(BX* BY) k_pressure_update_scalar(
    Grid g, double* __restrict__ p, const float* __restrict__ ut, const float* __restrict__ vt,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid,
    float* __restrict__ uo, float* __restrict__ vo, float dt, float k, DevScalars* sc,
    int own0, int own1) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  float ru = 0.f, rv = 0.f;
  const bool inside = x < g.nx && y < g.ny;
  if (inside) {
    const size_t c = gidx(g, x, y);
    const double pc = clamped_p(p, fluid, c);
    if (x < g.nx - 1 && !face_has<FACE_U>(solid, g, x, y) && face_has<FACE_U>(fluid, g, x, y)) {
      const float dp = (float)(clamped_p(p, fluid, c + 1) - pc);     // main.c:787, 705-707
      ru = ut[c] + (-k * dp) * dt;
    }
    if (y + g.yoff < g.gny - 1 && !face_has<FACE_V>(solid, g, x, y) && face_has<FACE_V>(fluid, g, x, y)) {
      const float dp = (float)(clamped_p(p, fluid, c + g.pitch) - pc);  // main.c:800
      rv = vt[c] + (-k * dp) * dt;
    }
    uo[c] = ru;
    vo[c] = rv;
  }
  // p is clamped in place only after every thread of the grid could have read the
  // unclamped neighbour: the clamp is idempotent, so writing it here is race-free in value
  // (a neighbour reads either p<0 and clamps it itself, or the already clamped 0).
  if (inside) {
    const size_t c = gidx(g, x, y);
    if (fluid[c] && p[c] < 0.0) p[c] = 0.0;
  }
  // fused max u^2 / max v^2 for the next calculate_timestep (main.c:808-820)
  const bool owned = y >= own0 && y < own1;
  float mu = ru * ru, mv = rv * rv;
  mu = (owned && mu > 0.f) ? mu : 0.f;     // drops NaN like the reference's `value > max`
  mv = (owned && mv > 0.f) ? mv : 0.f;
  mu = warp_maxf(mu);
  mv = warp_maxf(mv);
  if ((threadIdx.x & 31) == 0) {
    if (mu > 0.f) atomicMax(&sc->max_u2_bits, __float_as_uint(mu));
    if (mv > 0.f) atomicMax(&sc->max_v2_bits, __float_as_uint(mv));
  }
}


// ============================================================== 4 cells per thread ====

constexpr int QX = 32, QY = 8;           // threads per block: 32 quads (128 cells) x 8 rows
// persistent launch of the 4-cells-per-thread kernels: enough blocks to fill the machine, never
// more than there are pieces
inline int grid4(const Ctx& c) {
  const long pieces = (long)c.gt_tx * c.gt_ty * GT_SUB, want = (long)c.sm_count * 8;
  return (int)(pieces < want ? pieces : want);
}

// The fluid plane, read by every quad: ALL 32 lanes of the warp must call this (the byte right
// of the quad is the next lane's first byte; lane 31 loads it).
__device__ __forceinline__ QuadMask load_quad_mask_warp(const uint8_t* __restrict__ m, const Grid& g, size_t c) {
  QuadMask q;
  q.c = ld_u8x4(m + c);
  q.up = ld_u8x4(m + c + g.pitch);
  q.r = __shfl_down_sync(EULER_FULL_MASK, q.c, 1);
  if ((threadIdx.x & 31) == 31) q.r = m[c + 4];
  return q;
}
// Quad addressing shared by the four kernels.  A warp is one row of 32 quads; lanes past the
// row end keep a clamped, harmless address so that they can take part in the shuffle.
struct QuadPos { int x0, y; size_t c; bool row_ok, inside; };
__device__ __forceinline__ QuadPos quad_pos(const Grid& g, int bx, int by) {
  QuadPos q;
  q.x0 = (bx * QX + threadIdx.x) * 4;
  q.y = by * QY + threadIdx.y;
  q.row_ok = q.y < g.ny;                         // uniform per warp
  q.inside = q.row_ok && q.x0 < g.pitch;
  q.c = gidx(g, q.x0 < g.pitch ? q.x0 : g.pitch - 4, q.row_ok ? q.y : 0);
  return q;
}

// The kernels below are persistent: a block walks the 128-cell x 8-row pieces (16 per tile) of the
// tiles in the list, `body(bx, by)` once per piece with the piece's block coordinates.
static_assert(GT_W == 4 * QX * 4 && GT_H == 4 * QY && GT_SUB == 16, "a tile is 4 x 4 blocks");
template <class Body>
__device__ __forceinline__ void for_each_piece(const GridTiles& gt, Body body) {
  const unsigned int n = gt.list ? *gt.count : (unsigned int)(gt.tx * gt.ty);
  for (unsigned int i = blockIdx.x; i < n * GT_SUB; i += gridDim.x) {
    const int tile = gt.list ? gt.list[i / GT_SUB] : (int)(i / GT_SUB);
    const int sub = (int)(i % GT_SUB);
    body((tile % gt.tx) * 4 + (sub & 3), (tile / gt.tx) * 4 + (sub >> 2));
  }
}

// ---- which tiles the grid stages have to stream -----------------------------------------------
// k_gt_flags: per tile "holds fluid now".  k_gt_compact: a tile is streamed when it or one of its 8
// neighbours held fluid in this or one of the two previous sub-steps — every value a stage could
// produce or would have to clear lies within one cell of such a tile (faces touching fluid,
// main.c:128-138, 827), and planes written one or two sub-steps ago (the u/uext ping-pong) are
// still cleared where the fluid has left; everywhere else every plane is, and stays, zero.
__global__ void __launch_bounds__(128) k_gt_flags(Grid g, int tx, int ty, const uint8_t* __restrict__ fluid,
                                                   uint8_t* __restrict__ flags) {
  for (int tile = blockIdx.x; tile < tx * ty; tile += gridDim.x) {
    const int x0 = (tile % tx) * GT_W + threadIdx.x * 4, y0 = (tile / tx) * GT_H;
    const int y1 = min(y0 + GT_H, g.ny);
    unsigned any = 0;
    if (x0 < g.pitch)
      for (int y = y0; y < y1; ++y) any |= ld_u8x4(fluid + gidx(g, x0, y));
    const int has = __syncthreads_or(any != 0);
    if (threadIdx.x == 0) flags[tile] = has ? 1 : 0;
  }
}

__global__ void __launch_bounds__(1024) k_gt_compact(int tx, int ty, const uint8_t* __restrict__ f0,
                                                     const uint8_t* __restrict__ f1, const uint8_t* __restrict__ f2,
                                                     int* __restrict__ list, DevScalars* sc) {
  __shared__ int sh[1024];
  const int n = tx * ty;
  const int per = (n + 1023) / 1024;
  const int lo = min(n, per * (int)threadIdx.x), hi = min(n, lo + per);
  auto wanted = [&](int t) {
    const int x = t % tx, y = t / tx;
    for (int yy = max(y - 1, 0); yy <= min(y + 1, ty - 1); ++yy)
      for (int xx = max(x - 1, 0); xx <= min(x + 1, tx - 1); ++xx) {
        const int k = yy * tx + xx;
        if (f0[k] | f1[k] | f2[k]) return true;
      }
    return false;
  };
  int cnt = 0;
  for (int i = lo; i < hi; ++i) cnt += wanted(i) ? 1 : 0;
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
    if (wanted(i)) list[run++] = i;
  if (threadIdx.x == 1023) sc->grid_tiles = (unsigned int)sh[1023];
}

// Out of place: a face that stops being wet is zeroed here while a neighbour may still need
// its old value for the 3x3 mean (in the reference the two passes are sequential).
__global__ void __launch_bounds__(QX* QY) k_extrapolate_bounds(
    Grid g, GridTiles gt, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ prev,
    const uint8_t* __restrict__ solid, float* __restrict__ uo, float* __restrict__ vo) {
  for_each_piece(gt, [&](int bx, int by) {
    const QuadPos q = quad_pos(g, bx, by);
    if (!q.row_ok) return;
    const QuadMask f = load_quad_mask_warp(fluid, g, q.c);
    if (!q.inside) return;
    F4 ru, rv;
#pragma unroll
    for (int k = 0; k < 4; ++k) ru.v[k] = rv.v[k] = 0.f;
    if (f.any()) extrapolate_quad(g, f, q.x0, q.y, q.c, u, v, prev, solid, ru, rv);
    st_f4(uo + q.c, ru);
    st_f4(vo + q.c, rv);
  });
}

// ---- advect_u, advect_v + gravity + zero_bounds ---------------------------------------
template <int MINB>
__global__ void __launch_bounds__(QX* QY, MINB) k_advect_velocity(
    Grid g, GridTiles gt, InterpLimits lim, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid,
    float* __restrict__ uo, float* __restrict__ vo, float dt, float h, float gravity) {
  for_each_piece(gt, [&](int bx, int by) {
    const QuadPos q = quad_pos(g, bx, by);
    if (!q.row_ok) return;
    const QuadMask f = load_quad_mask_warp(fluid, g, q.c);
    if (!q.inside) return;
    F4 ru, rv;
#pragma unroll
    for (int k = 0; k < 4; ++k) ru.v[k] = rv.v[k] = 0.f;
    if (f.any()) advect_quad(g, lim, f, q.x0, q.y, q.c, u, v, fluid, solid, dt, h, gravity, ru, rv);
    st_f4(uo + q.c, ru);
    st_f4(vo + q.c, rv);
  });
}

// ---- rhs build --------------------------------------------------------------------------
__global__ void __launch_bounds__(QX* QY) k_build_rhs(
    Grid g, GridTiles gt, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid,
    double* __restrict__ r, double* __restrict__ p, int8_t* __restrict__ adiag, float h,
    double scale, DevScalars* sc, int own0, int own1, float* __restrict__ r32) {
  bool nz = false;
  for_each_piece(gt, [&](int bx, int by) {
    const QuadPos q = quad_pos(g, bx, by);
    if (!q.inside) return;
    const unsigned mf = ld_u8x4(fluid + q.c);
    D4g b, zero;
#pragma unroll
    for (int k = 0; k < 4; ++k) b.v[k] = zero.v[k] = 0.0;
    if (mf) {
      const bool owned = q.y >= own0 && q.y < own1;      // halo rows are the neighbour slab's business
      nz |= rhs_quad(g, mf, q.c, u, v, solid, adiag, h, scale, owned, b);
    }
    st_d4(r + q.c, b);
    st_d4(p + q.c, zero);                                 // p = 0, main.c:739
    if (r32)                                              // mixed-precision PCG: r starts as fp32(b)
      *reinterpret_cast<float4*>(r32 + q.c) = make_float4((float)b.v[0], (float)b.v[1], (float)b.v[2], (float)b.v[3]);
  });
  if (__any_sync(EULER_FULL_MASK, nz) && (threadIdx.x & 31) == 0) atomicOr(&sc->nonzero_rhs, 1);
}

// ---- pressure update ----------------------------------------------------------------------
__global__ void __launch_bounds__(QX* QY) k_pressure_update(
    Grid g, GridTiles gt, double* __restrict__ p, const float* __restrict__ ut, const float* __restrict__ vt,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid,
    float* __restrict__ uo, float* __restrict__ vo, float dt, float kk, DevScalars* sc,
    int own0, int own1) {
  float mu = 0.f, mv = 0.f;
  for_each_piece(gt, [&](int bx, int by) {
    const QuadPos q = quad_pos(g, bx, by);
    if (!q.row_ok) return;
    const QuadMask f = load_quad_mask_warp(fluid, g, q.c);
    if (!q.inside) return;
    F4 ru, rv;
#pragma unroll
    for (int k = 0; k < 4; ++k) ru.v[k] = rv.v[k] = 0.f;
    if (f.any()) pressure_quad(g, f, q.x0, q.y, q.c, p, ut, vt, solid, dt, kk, ru, rv);
    st_f4(uo + q.c, ru);
    st_f4(vo + q.c, rv);
    // fused max u^2 / max v^2 for the next calculate_timestep (main.c:808-820); cells outside
    // the streamed tiles are zero and cannot raise a maximum
    if (q.y >= own0 && q.y < own1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = ru.v[k] * ru.v[k], b = rv.v[k] * rv.v[k];
        if (a > mu) mu = a;                    // drops NaN like the reference's `value > max`
        if (b > mv) mv = b;
      }
    }
  });
  mu = warp_maxf(mu);
  mv = warp_maxf(mv);
  if ((threadIdx.x & 31) == 0) {
    if (mu > 0.f) atomicMax(&sc->max_u2_bits, __float_as_uint(mu));
    if (mv > 0.f) atomicMax(&sc->max_v2_bits, __float_as_uint(mv));
  }
}


// ===================================================== --rainbow colour transport ====
// A passive RGB scalar on the P cells (reference main.c:76-84): colorize :187-201,
// extrapolate(P) :859-863, source colours :292-294, advect_p :424-438 + the whole-plane copies
// :873-882.  Off unless euler_params.rainbow; single-GPU handles only.

// misc/color.h hsv_basis: periodic in t with period 6, values in [0,1]
__device__ __forceinline__ float hsv_basis(float t) {
  t -= 6.f * floorf(1.f / 6 * t);
  if (t < 0.f) t += 6.f;
  if (t < 1.f) return t;
  if (t < 3.f) return 1.f;
  if (t < 4.f) return 4.f - t;
  return 0.f;
}

// colorize(), main.c:187-201: hue ramps along x+y with a period of 60 cells (k_initial_color_
// period); source cells start at t = 0.  Fluid cells only.
__global__ void __launch_bounds__(BX* BY) k_colorize(
    Grid g, const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ source,
    float* __restrict__ cr, float* __restrict__ cg, float* __restrict__ cb) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  if (x >= g.nx || y >= g.ny) return;
  const size_t c = gidx(g, x, y);
  if (!fluid[c]) return;
  float t = 0.f;
  if (!source[c]) t = (x + y + g.yoff) * 6.f / 60.f;
  cr[c] = hsv_basis(t + 2.f);
  cg[c] = hsv_basis(t);
  cb[c] = hsv_basis(t - 2.f);
}

// extrapolate(g_r|g_g|g_b, P), main.c:859-863 with :158-185: a cell that is fluid now but was
// not last sub-step takes the mean of the cells of its clamped 3x3 block that were.  In place:
// the cells written are exactly those no other cell reads (they fail the `prev` test).
__global__ void __launch_bounds__(BX* BY) k_extrapolate_color(
    Grid g, const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ prev,
    float* __restrict__ cr, float* __restrict__ cg, float* __restrict__ cb) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  if (x >= g.nx || y >= g.ny) return;
  const size_t c = gidx(g, x, y);
  if (prev[c] || !fluid[c]) return;
  const int x0 = max(x - 1, 0), x1 = min(x + 1, g.nx - 1);
  const int y0 = max(y - 1, -g.yoff), y1 = min(y + 1, g.gny - 1 - g.yoff);
  float tr = 0.f, tg = 0.f, tb = 0.f;
  int n = 0;
  for (int yy = y0; yy <= y1; ++yy)
    for (int xx = x0; xx <= x1; ++xx) {
      const size_t k = gidx(g, xx, yy);
      if (prev[k]) { tr += cr[k]; tg += cg[k]; tb += cb[k]; ++n; }
    }
  cr[c] = tr / (float)n;            // n == 0 -> NaN, as in the reference (assert off)
  cg[c] = tg / (float)n;
  cb[c] = tb / (float)n;
}

// the colour writes of update_fluid_sources, main.c:283, 292-294
__global__ void k_source_colors(const unsigned int* __restrict__ cells, size_t n, float t,
                                float* __restrict__ cr, float* __restrict__ cg, float* __restrict__ cb) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned int c = cells[i];
  cr[c] = hsv_basis(t + 2.f);
  cg[c] = hsv_basis(t);
  cb[c] = hsv_basis(t - 2.f);
}

// advect_p x 3, main.c:424-438: back-trace from the cell centre with the mean of the two faces
// either side, masked-bilinear sample of each colour plane there.  Only fluid cells are
// written (the tmp planes keep their own older values elsewhere, like the reference's).
__global__ void __launch_bounds__(BX* BY) k_advect_color(
    Grid g, InterpLimits lim, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const float* __restrict__ cr, const float* __restrict__ cg,
    const float* __restrict__ cb, float* __restrict__ ro, float* __restrict__ go,
    float* __restrict__ bo, float dt, float h) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  if (x >= g.nx || y >= g.ny) return;
  const size_t c = gidx(g, x, y);
  if (!fluid[c]) return;
  const float dy = (v[c] + v[c - g.pitch]) / 2;
  const float dx = (u[c] + u[c - 1]) / 2;
  const float px = x - div_h(dx * dt, h);
  const float py = (y + g.yoff) - div_h(dy * dt, h);
  ro[c] = interpolate<CELL_P>(cr, fluid, g, lim, px, py);
  go[c] = interpolate<CELL_P>(cg, fluid, g, lim, px, py);
  bo[c] = interpolate<CELL_P>(cb, fluid, g, lim, px, py);
}

// ------------------------------------------------------------------ invariants ----
// euler_gpu_check: size-independent invariants of the state over the OWNED rows (one pass).
// Integer results by 64-bit atomics (exact, order-free); floating sums as one partial per block,
// added up by the host in block order (deterministic for a given grid).
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

__global__ void __launch_bounds__(256) k_check(Grid g, int r0, int r1, const float* __restrict__ u,
                                                const float* __restrict__ v, const double* __restrict__ p,
                                                const uint8_t* __restrict__ count, float h,
                                                unsigned long long* __restrict__ ints /*3*/,
                                                double* __restrict__ parts /*6 x gridDim.x*/) {
  unsigned long long n_fluid = 0, c_sum = 0, c_hash = 0;
  double su = 0.0, sv = 0.0, sp = 0.0, mdiv = 0.0, mu = 0.0, mv = 0.0;
  const size_t total = (size_t)g.nx * (r1 - r0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int y = r0 + (int)(i / g.nx), x = (int)(i % g.nx);
    const size_t c = gidx(g, x, y);
    const unsigned cnt = count[c];
    const double au = fabs((double)u[c]), av = fabs((double)v[c]);
    if (au == au) { su += au; mu = fmax(mu, au); }
    if (av == av) { sv += av; mv = fmax(mv, av); }
    if (cnt) {
      n_fluid += 1; c_sum += cnt;
      c_hash += cnt * mix64((unsigned long long)(y + g.yoff) * (unsigned long long)g.nx + x + 1);
      if (p) sp += p[c];
      const float div = div_h(u[c] - u[c - 1] + v[c] - v[c - g.pitch], h);   // main.c:720
      const double ad = fabs((double)div);
      if (ad == ad) mdiv = fmax(mdiv, ad);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    n_fluid += __shfl_xor_sync(EULER_FULL_MASK, n_fluid, o);
    c_sum += __shfl_xor_sync(EULER_FULL_MASK, c_sum, o);
    c_hash += __shfl_xor_sync(EULER_FULL_MASK, c_hash, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(ints + 0, n_fluid); atomicAdd(ints + 1, c_sum); atomicAdd(ints + 2, c_hash); }
  double* out = parts + (size_t)blockIdx.x * 6;
  double r;
  r = block_reduce<false>(su); if (threadIdx.x == 0) out[0] = r;
  r = block_reduce<false>(sv); if (threadIdx.x == 0) out[1] = r;
  r = block_reduce<false>(sp); if (threadIdx.x == 0) out[2] = r;
  r = block_reduce<true>(mdiv); if (threadIdx.x == 0) out[3] = r;
  r = block_reduce<true>(mu); if (threadIdx.x == 0) out[4] = r;
  r = block_reduce<true>(mv); if (threadIdx.x == 0) out[5] = r;
}

}  // namespace

// ----------------------------------------------------------------- launchers ----

void launch_check(Ctx& c, unsigned long long* ints3, double* parts, int blocks) {
  cudaMemsetAsync(ints3, 0, 3 * sizeof(unsigned long long), c.stream);
  k_check<<<blocks, 256, 0, c.stream>>>(c.g, c.own0, c.own1, c.u, c.v, c.p, c.count, c.h, ints3, parts);
  c.launches += 1;
}

// The tile list of this sub-step, from the final count plane (after the sources and, on a slab,
// the halo exchange of the classification).  Until two earlier flag planes exist the stages run
// over all tiles (and thereby bring every plane into the "zero outside the list" state).
void launch_grid_tiles(Ctx& c) {
  ProfScope ps(c, KC_MISC);
  uint8_t* now = c.gt_flags[2];                 // rotate: [0] = now, [1], [2] = the two previous
  c.gt_flags[2] = c.gt_flags[1]; c.gt_flags[1] = c.gt_flags[0]; c.gt_flags[0] = now;
  const int n = c.gt_tx * c.gt_ty;
  k_gt_flags<<<n < c.sm_count * 16 ? n : c.sm_count * 16, 128, 0, c.stream>>>(c.g, c.gt_tx, c.gt_ty, c.count, now);
  c.launches += 1;
  static const int off = getenv("EULER_GRID_DENSE") ? atoi(getenv("EULER_GRID_DENSE")) : 0;   // A/B knob
  c.gt_sparse = (c.gt_hist >= 2 && !off) ? 1 : 0;
  if (c.gt_sparse) {
    k_gt_compact<<<1, 1024, 0, c.stream>>>(c.gt_tx, c.gt_ty, c.gt_flags[0], c.gt_flags[1], c.gt_flags[2], c.gt_list, c.sc);
    c.launches += 1;
  }
  if (c.gt_hist < 2) c.gt_hist++;
}

void launch_maxsq(Ctx& c) {
  ProfScope ps(c, KC_MAXSQ);
  cudaMemsetAsync(&c.sc->max_u2_bits, 0, 2 * sizeof(unsigned int), c.stream);
  const int blocks = c.sm_count * 8;
  k_maxsq<<<blocks, 256, 0, c.stream>>>(c.g, c.own0, c.own1, c.u, c.v, c.sc);
  c.launches += 1;
}

void launch_timestep(Ctx& c, float frame_time, float cfl) {
  ProfScope ps(c, KC_MISC);
  k_timestep<<<1, 1, 0, c.stream>>>(c.sc, cfl * c.h, frame_time);
  c.launches += 1;
}

// EULER_GRID_VARIANT=1 selects the one-cell-per-thread kernels (A/B runs)
static int grid_variant() {
  static const int v = getenv("EULER_GRID_VARIANT") ? atoi(getenv("EULER_GRID_VARIANT")) : 0;
  return v;
}
static bool scalar_variant() { return grid_variant() == 1; }

void launch_extrapolate(Ctx& c) {
  ProfScope ps(c, KC_EXTRAPOLATE);
  if (scalar_variant())
    k_extrapolate_bounds_scalar<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(
        c.g, c.u, c.v, c.count, c.prev_count, c.solid, c.uext, c.vext);
  else
    k_extrapolate_bounds<<<grid4(c), dim3(QX, QY), 0, c.stream>>>(
        c.g, grid_tiles_of(c, c.gt_sparse), c.u, c.v, c.count, c.prev_count, c.solid, c.uext, c.vext);
  c.launches += 1;
}

void launch_advect_velocity(Ctx& c, float dt) {
  ProfScope ps(c, KC_ADVECT_VELOCITY);
  if (scalar_variant())
    k_advect_velocity_scalar<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(
        c.g, c.lim, c.u, c.v, c.count, c.solid, c.utmp, c.vtmp, dt, c.h, c.gravity);
  else {
    // Resident blocks per SM the compiler must allow: the kernel waits on dependent gathers
    // (ncu: 35 % long-scoreboard stalls, sm throughput 69 %), so more warps beat more registers.
    // Measured at 16384^2: 1.21 ms at 4-5 blocks (48-56 registers), 1.17 ms at 6 (40), 1.12 ms
    // at 8 (32 registers, 48 B of spills).  Taking the two fixed-offset samples from registers
    // (quad-wide mask/value loads instead of 10 loads per sample) was tried as well: same time at
    // equal occupancy, so the simpler kernel stays.
    static const int minb = getenv("EULER_ADV_MINB") ? atoi(getenv("EULER_ADV_MINB")) : 8;
#define ADV(M) k_advect_velocity<M><<<grid4(c), dim3(QX, QY), 0, c.stream>>>( \
        c.g, grid_tiles_of(c, c.gt_sparse), c.lim, c.u, c.v, c.count, c.solid, c.utmp, c.vtmp, dt, c.h, c.gravity)
    if (minb == 4) ADV(4); else if (minb == 6) ADV(6); else ADV(8);
#undef ADV
  }
  c.launches += 1;
}

void launch_build_rhs(Ctx& c, float dt) {
  ProfScope ps(c, KC_BUILD_RHS);
  cudaMemsetAsync(&c.sc->nonzero_rhs, 0, sizeof(int), c.stream);
  const double scale = (double)((c.h * c.h) * c.rho / dt);     // fp32 expression, main.c:713
  if (scalar_variant())
    k_build_rhs_scalar<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(
        c.g, c.utmp, c.vtmp, c.count, c.solid, c.r, c.p, c.adiag, c.h, scale, c.sc, c.own0, c.own1,
        c.mixed ? c.r32 : nullptr);
  else
    k_build_rhs<<<grid4(c), dim3(QX, QY), 0, c.stream>>>(
        c.g, grid_tiles_of(c, c.gt_sparse), c.utmp, c.vtmp, c.count, c.solid, c.r, c.p, c.adiag, c.h, scale, c.sc,
        c.own0, c.own1, c.mixed ? c.r32 : nullptr);
  c.launches += 1;
}

void launch_pressure_update(Ctx& c, float dt) {
  ProfScope ps(c, KC_PRESSURE_UPDATE);
  cudaMemsetAsync(&c.sc->max_u2_bits, 0, 2 * sizeof(unsigned int), c.stream);
  const float k = 1.f / (c.rho * c.h);                          // invf(rho*h), main.c:706
  if (scalar_variant())
    k_pressure_update_scalar<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(
        c.g, c.p, c.utmp, c.vtmp, c.count, c.solid, c.u, c.v, dt, k, c.sc, c.own0, c.own1);
  else
    k_pressure_update<<<grid4(c), dim3(QX, QY), 0, c.stream>>>(
        c.g, grid_tiles_of(c, c.gt_sparse), c.p, c.utmp, c.vtmp, c.count, c.solid, c.u, c.v, dt, k, c.sc, c.own0, c.own1);
  c.launches += 1;
}

// ---- --rainbow (no-ops unless the colour planes exist) -----------------------------------
void launch_colorize(Ctx& c) {
  if (!c.cr) return;
  ProfScope ps(c, KC_COLOR);
  k_colorize<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(c.g, c.count, c.source, c.cr, c.cg, c.cb);
  c.launches += 1;
}

void launch_extrapolate_color(Ctx& c) {
  if (!c.cr) return;
  ProfScope ps(c, KC_COLOR);
  k_extrapolate_color<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(c.g, c.count, c.prev_count, c.cr, c.cg, c.cb);
  c.launches += 1;
}

void launch_source_colors(Ctx& c, unsigned int frame_count) {
  if (!c.cr || !c.n_source_cells) return;
  ProfScope ps(c, KC_COLOR);
  const float t = 0.6f / 10.f * (float)(unsigned short)frame_count;    // main.c:83, 283 (uint16 count)
  k_source_colors<<<(unsigned)((c.n_source_cells + 255) / 256), 256, 0, c.stream>>>(
      c.source_cells, c.n_source_cells, t, c.cr, c.cg, c.cb);
  c.launches += 1;
}

void launch_advect_color(Ctx& c, float dt) {
  if (!c.cr) return;
  ProfScope ps(c, KC_COLOR);
  k_advect_color<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(
      c.g, c.lim, c.u, c.v, c.count, c.cr, c.cg, c.cb, c.crtmp, c.cgtmp, c.cbtmp, dt, c.h);
  c.launches += 1;
  // memcpy(g_r, g_rtmp, sizeof(g_r)) etc., main.c:875-881: whole planes
  const size_t bytes = (size_t)c.g.ny * c.g.pitch * sizeof(float);
  cudaMemcpyAsync(c.cr, c.crtmp, bytes, cudaMemcpyDeviceToDevice, c.stream);
  cudaMemcpyAsync(c.cg, c.cgtmp, bytes, cudaMemcpyDeviceToDevice, c.stream);
  cudaMemcpyAsync(c.cb, c.cbtmp, bytes, cudaMemcpyDeviceToDevice, c.stream);
}

}  // namespace euler
