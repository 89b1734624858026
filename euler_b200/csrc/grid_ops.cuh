// euler_b200/csrc/grid_ops.cuh — what the MAC-grid stage kernels (grid_kernels.cu) compute for ONE
// quad of four consecutive cells: extrapolate + zero_bounds, advect_u/v + gravity + bounds, rhs +
// a_diag, pressure clamp + gradient.  The kernels add the addressing, the warp-level mask exchange
// and the reductions; the arithmetic lives here so that the same source also compiles for the
// host, where tests/test_kernel_arith_host.py checks it bit for bit against the oracle without a GPU.
// Reference lines: main.c:158-185, 382-422, 539-545, 713-733, 769-805, 822-832.
#pragma once
#include <math.h>

#include "common.cuh"
#include "interp.cuh"

#if !defined(__CUDACC__) && !defined(__noinline__)
#define __noinline__ __attribute__((noinline))      // host build: nvcc's keyword
#endif

namespace euler {

namespace {

__device__ __forceinline__ unsigned ld_u8x4(const uint8_t* __restrict__ p) {
  return *reinterpret_cast<const unsigned*>(p);
}
__device__ __forceinline__ bool byte_set(unsigned m, int k) { return ((m >> (8 * k)) & 0xffu) != 0; }
__device__ __forceinline__ int byte_of(unsigned m, int k) { return (int)((m >> (8 * k)) & 0xffu); }

// Masks of one quad's neighbourhood in one u8 plane: the quad itself, the cell right of it and
// the row above — everything the U faces (x, x+1) and V faces (y, y+1) of four cells need.
struct QuadMask {
  unsigned c, up;     // 4 bytes each
  unsigned r;         // low byte: cell x0+4 of the quad's row
  __device__ __forceinline__ bool cell(int k) const { return byte_set(c, k); }
  __device__ __forceinline__ bool right(int k) const { return k == 3 ? (r & 0xffu) != 0 : byte_set(c, k + 1); }
  __device__ __forceinline__ bool above(int k) const { return byte_set(up, k); }
  __device__ __forceinline__ bool face_u(int k) const { return cell(k) | right(k); }   // main.c:128-132
  __device__ __forceinline__ bool face_v(int k) const { return cell(k) | above(k); }   // main.c:134-138
  __device__ __forceinline__ bool any() const { return (c | up | (r & 0xffu)) != 0; }
};
// Any other plane, read only by the quads that touch fluid (divergent code: no shuffle).
__device__ __forceinline__ QuadMask load_quad_mask(const uint8_t* __restrict__ m, const Grid& g, size_t c) {
  QuadMask q;
  q.c = ld_u8x4(m + c);
  q.up = ld_u8x4(m + c + g.pitch);
  q.r = m[c + 4];
  return q;
}

struct F4 { float v[4]; };
__device__ __forceinline__ F4 ld_f4(const float* __restrict__ p) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  F4 r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  return r;
}
__device__ __forceinline__ void st_f4(float* __restrict__ p, const F4& a) {
  *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
struct D4g { double v[4]; };
__device__ __forceinline__ D4g ld_d4(const double* __restrict__ p) {
  const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  D4g r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
  return r;
}
__device__ __forceinline__ void st_d4(double* __restrict__ p, const D4g& a) {
  *reinterpret_cast<double2*>(p) = make_double2(a.v[0], a.v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(a.v[2], a.v[3]);
}

// ---- extrapolate + zero_bounds ------------------------------------------------------
// mean of the clamped 3x3 block's faces that were wet last sub-step (main.c:158-171, 179-181),
// row-major accumulation order; 0/0 -> NaN when there is none (the assert is off).  The rare
// path of extrapolate: a face that has just become wet.
template <int TYPE>
__device__ __noinline__ float newly_wet_mean(const Grid& g, const float* __restrict__ q,
                                             const uint8_t* __restrict__ prev, int x, int y) {
  const int sx = g.nx - (TYPE == FACE_U), sy = g.gny - (TYPE == FACE_V) - g.yoff;
  const int x0 = max(x - 1, 0), x1 = min(x + 1, sx - 1);
  const int y0 = max(y - 1, -g.yoff), y1 = min(y + 1, sy - 1);
  float total = 0.f;
  int n = 0;
  for (int yy = y0; yy <= y1; ++yy)
    for (int xx = x0; xx <= x1; ++xx)
      if (face_has<TYPE>(prev, g, xx, yy)) { total += q[gidx(g, xx, yy)]; ++n; }
  return total / (float)n;
}

// ---- one quad of extrapolate + zero_bounds (u and v), main.c:173-185, 822-832 ----------------
__device__ __forceinline__ void extrapolate_quad(const Grid& g, const QuadMask& f, int x0, int y, size_t c,
                                                 const float* __restrict__ u, const float* __restrict__ v,
                                                 const uint8_t* __restrict__ prev, const uint8_t* __restrict__ solid,
                                                 F4& ru, F4& rv) {
  const QuadMask s = load_quad_mask(solid, g, c), pv = load_quad_mask(prev, g, c);
  const F4 uc = ld_f4(u + c), vc = ld_f4(v + c);
  const bool v_row = y + g.yoff < g.gny - 1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x0 + k;
    // zero_bounds (main.c:827): not touching fluid, or touching a solid -> 0
    if (x < g.nx - 1 && f.face_u(k) && !s.face_u(k))
      ru.v[k] = pv.face_u(k) ? uc.v[k] : newly_wet_mean<FACE_U>(g, u, prev, x, y);
    if (x < g.nx && v_row && f.face_v(k) && !s.face_v(k))
      rv.v[k] = pv.face_v(k) ? vc.v[k] : newly_wet_mean<FACE_V>(g, v, prev, x, y);
  }
}

// ---- one quad of advect_u, advect_v + gravity + zero_bounds, main.c:382-422, 539-545, 888-889 --
__device__ __forceinline__ void advect_quad(const Grid& g, const InterpLimits& lim, const QuadMask& f, int x0, int y,
                                            size_t c, const float* __restrict__ u, const float* __restrict__ v,
                                            const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid,
                                            float dt, float h, float gravity, F4& ru, F4& rv) {
  const QuadMask s = load_quad_mask(solid, g, c);
  const F4 uc = ld_f4(u + c), vc = ld_f4(v + c);
  const int gy = y + g.yoff;               // sample positions are in global index space
  const bool v_row = gy < g.gny - 1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x0 + k;
    if (x < g.nx - 1 && f.face_u(k) && !s.face_u(k)) {
      // main.c:388-395: back-trace one Euler step, sample u there
      const float dx = uc.v[k];
      const float dy = interpolate<FACE_V>(v, fluid, g, lim, x + 0.5f, gy - 0.5f);
      const float px = x - div_h(dx * dt, h);
      const float py = gy - div_h(dy * dt, h);
      ru.v[k] = interpolate<FACE_U>(u, fluid, g, lim, px, py);
    }
    if (x < g.nx && v_row && f.face_v(k) && !s.face_v(k)) {
      // main.c:411-418, then gravity main.c:542
      const float dy = vc.v[k];
      const float dx = interpolate<FACE_U>(u, fluid, g, lim, x - 0.5f, gy + 0.5f);
      const float px = x - div_h(dx * dt, h);
      const float py = gy - div_h(dy * dt, h);
      float r = interpolate<FACE_V>(v, fluid, g, lim, px, py);
      r += gravity * dt;
      rv.v[k] = r;
    }
  }
}

// ---- one quad of the rhs and a_diag, main.c:713-733; returns "some owned b is nonzero" ---------
__device__ __forceinline__ bool rhs_quad(const Grid& g, unsigned mf, size_t c, const float* __restrict__ u,
                                         const float* __restrict__ v, const uint8_t* __restrict__ solid,
                                         int8_t* __restrict__ adiag, float h, double scale, bool owned, D4g& b) {
  bool nz = false;
  const F4 uc = ld_f4(u + c), vc = ld_f4(v + c), vd = ld_f4(v + c - g.pitch);
  const float ul = u[c - 1];
  const uint8_t* sp = solid + c;
  const unsigned s_c = ld_u8x4(sp), s_dn = ld_u8x4(sp - g.pitch), s_up = ld_u8x4(sp + g.pitch);
  const int s_l = sp[-1], s_r = sp[4];
  unsigned am = ld_u8x4(reinterpret_cast<const uint8_t*>(adiag) + c);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!byte_set(mf, k)) continue;
    // main.c:720-721: divergence left to right in fp32, widened, scaled by h^2 rho/dt
    const float div = div_h(uc.v[k] - (k == 0 ? ul : uc.v[k - 1]) + vc.v[k] - vd.v[k], h);
    b.v[k] = -(double)div * scale;
    // main.c:554-559: 4 minus the number of solid neighbours (left, right, down, up)
    const int a = 4 - (k == 0 ? s_l : byte_of(s_c, k - 1)) - (k == 3 ? s_r : byte_of(s_c, k + 1)) -
                  byte_of(s_dn, k) - byte_of(s_up, k);
    am = (am & ~(0xffu << (8 * k))) | (((unsigned)a & 0xffu) << (8 * k));
    nz |= (b.v[k] != 0.0) && owned;
  }
  *reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(adiag) + c) = am;
  return nz;
}

// ---- one quad of the pressure clamp + gradient update, main.c:769-805 ------------------------
__device__ __forceinline__ void pressure_quad(const Grid& g, const QuadMask& f, int x0, int y, size_t c,
                                              double* __restrict__ p, const float* __restrict__ ut,
                                              const float* __restrict__ vt, const uint8_t* __restrict__ solid,
                                              float dt, float kk, F4& ru, F4& rv) {
  const QuadMask s = load_quad_mask(solid, g, c);
  D4g pc = ld_d4(p + c), pu = ld_d4(p + c + g.pitch);
  double pr = p[c + 4];
  const F4 uc = ld_f4(ut + c), vc = ld_f4(vt + c);
  // p = max(p, 0) on fluid cells (main.c:773-779), applied to every operand as it is read
  bool clamped = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (f.cell(k) && pc.v[k] < 0.0) { pc.v[k] = 0.0; clamped = true; }
    if (f.above(k) && pu.v[k] < 0.0) pu.v[k] = 0.0;
  }
  if ((f.r & 0xffu) && pr < 0.0) pr = 0.0;
  const bool v_row = y + g.yoff < g.gny - 1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x0 + k;
    if (x < g.nx - 1 && !s.face_u(k) && f.face_u(k)) {
      const float dp = (float)((k == 3 ? pr : pc.v[k + 1]) - pc.v[k]);   // main.c:787, 705-707
      ru.v[k] = uc.v[k] + (-kk * dp) * dt;
    }
    if (x < g.nx && v_row && !s.face_v(k) && f.face_v(k)) {
      const float dp = (float)(pu.v[k] - pc.v[k]);                       // main.c:800
      rv.v[k] = vc.v[k] + (-kk * dp) * dt;
    }
  }
  // the clamp is idempotent, so writing it while neighbours may still read the old value
  // is race-free in value (they clamp what they read themselves)
  if (clamped) st_d4(p + c, pc);
}

}  // namespace

}  // namespace euler
