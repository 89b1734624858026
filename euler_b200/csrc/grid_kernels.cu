// euler_b200/csrc/grid_kernels.cu — the MAC-grid (non-PCG) stages of the sub-step.
//
//   k_maxsq + k_timestep      calculate_timestep / maxsq          reference main.c:808-841
//   k_extrapolate_bounds      extrapolate(u),(v) + zero_bounds    main.c:158-185, 822-832, 865-868
//   k_advect_velocity         advect_u, advect_v, apply_body_forces, zero_bounds(tmp)
//                                                                 main.c:382-422, 539-545, 871-889
//   k_build_rhs               b, a_diag, p=0 (first part of project) main.c:713-733, 739
//   k_pressure_update         clamp p, subtract grad p            main.c:769-805
//
// All of them are HBM-bound streaming stencils (DESIGN.md §4 has the byte counts).  One
// thread owns one P cell and the U face on its right and the V face above it; a warp covers
// 32 consecutive x, so every plane access is a fully coalesced 32/128/256 B row segment.
#include "interp.cuh"
#include "kernels.h"

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
__global__ void __launch_bounds__(BX* BY) k_extrapolate_bounds(
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

__global__ void __launch_bounds__(BX* BY) k_advect_velocity(
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

__global__ void __launch_bounds__(BX* BY) k_build_rhs(
    Grid g, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid,
    double* __restrict__ r, double* __restrict__ p, int8_t* __restrict__ adiag, float h,
    double scale, DevScalars* sc, int own0, int own1) {
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
  }
  if (__any_sync(EULER_FULL_MASK, nz) && (threadIdx.x & 31) == 0) atomicOr(&sc->nonzero_rhs, 1);
}

// ------------------------------------------------------------ pressure update ----

__device__ __forceinline__ double clamped_p(const double* __restrict__ p,
                                            const uint8_t* __restrict__ fluid, size_t c) {
  double v = p[c];
  return (fluid[c] && v < 0.0) ? 0.0 : v;                    // main.c:773-779
}

__global__ void __launch_bounds__(BX* BY) k_pressure_update(
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

}  // namespace

// ----------------------------------------------------------------- launchers ----

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

void launch_extrapolate(Ctx& c) {
  ProfScope ps(c, KC_EXTRAPOLATE);
  k_extrapolate_bounds<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(
      c.g, c.u, c.v, c.count, c.prev_count, c.solid, c.uext, c.vext);
  c.launches += 1;
}

void launch_advect_velocity(Ctx& c, float dt) {
  ProfScope ps(c, KC_ADVECT_VELOCITY);
  k_advect_velocity<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(
      c.g, c.lim, c.u, c.v, c.count, c.solid, c.utmp, c.vtmp, dt, c.h, c.gravity);
  c.launches += 1;
}

void launch_build_rhs(Ctx& c, float dt) {
  ProfScope ps(c, KC_BUILD_RHS);
  cudaMemsetAsync(&c.sc->nonzero_rhs, 0, sizeof(int), c.stream);
  const double scale = (double)((c.h * c.h) * c.rho / dt);     // fp32 expression, main.c:713
  k_build_rhs<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(
      c.g, c.utmp, c.vtmp, c.count, c.solid, c.r, c.p, c.adiag, c.h, scale, c.sc, c.own0, c.own1);
  c.launches += 1;
}

void launch_pressure_update(Ctx& c, float dt) {
  ProfScope ps(c, KC_PRESSURE_UPDATE);
  cudaMemsetAsync(&c.sc->max_u2_bits, 0, 2 * sizeof(unsigned int), c.stream);
  const float k = 1.f / (c.rho * c.h);                          // invf(rho*h), main.c:706
  k_pressure_update<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(
      c.g, c.p, c.utmp, c.vtmp, c.count, c.solid, c.u, c.v, dt, k, c.sc, c.own0, c.own1);
  c.launches += 1;
}

}  // namespace euler
