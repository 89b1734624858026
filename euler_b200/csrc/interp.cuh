// euler_b200/csrc/interp.cuh — masked bilinear sampling of the staggered velocity planes.
//
// Follows reference main.c:301-364 (get_fraction, linear, bilinear, sparse_get, interpolate)
// operation for operation, in fp32 without FMA contraction, so that every sample is
// bit-identical to the reference's:
//   * the sample index is clamped to [0, nextafterf(size-1, 0)]            (main.c:339-340)
//   * split into whole and fractional part (modff)                          (main.c:342-344)
//   * a corner is usable iff it is a fluid face: a U face touches fluid if P(x) or P(x+1) is
//     fluid, a V face if P(y) or P(y+1) is                                  (main.c:128-138)
//   * unusable corners read as 0 and are excluded by snapping the fraction  (main.c:301-309)
//   * the two vertical lerps come first, then the horizontal one           (main.c:322-330)
#pragma once
#include "common.cuh"

namespace euler {

struct InterpLimits {   // nextafterf(size-1, 0) per plane, computed on the host with libm
  float u_x, u_y, v_x, v_y;
};

template <int TYPE>
__device__ __forceinline__ bool face_has(const uint8_t* __restrict__ m, const Grid& g, int x, int y) {
  const size_t c = gidx(g, x, y);
  bool a = m[c] != 0;
  if (TYPE == FACE_U) a |= (m[c + 1] != 0);
  if (TYPE == FACE_V) a |= (m[c + g.pitch] != 0);
  return a;
}

__device__ __forceinline__ float lerp_ref(float a, float b, float f) {
  return (1.f - f) * a + f * b;                              // main.c:311-313
}
__device__ __forceinline__ float snap_fraction(float f, bool lo_ok, bool hi_ok) {
  return !lo_ok ? 1.f : (!hi_ok ? 0.f : f);                  // main.c:301-309
}

template <int TYPE>
__device__ __forceinline__ float interpolate(const float* __restrict__ q,
                                             const uint8_t* __restrict__ fluid,
                                             const Grid& g, const InterpLimits& lim,
                                             float ix, float iy) {
  const float hx = TYPE == FACE_U ? lim.u_x : lim.v_x;
  const float hy = TYPE == FACE_U ? lim.u_y : lim.v_y;
  ix = ix < 0.f ? 0.f : (ix > hx ? hx : ix);
  iy = iy < 0.f ? 0.f : (iy > hy ? hy : iy);
  const float wx = truncf(ix), wy = truncf(iy);
  const float fx = ix - wx, fy = iy - wy;                    // == modff for finite input
  const int bx = (int)wx;
  // global row -> row of this view; clamped into the guard band so that a sample outside the
  // locally stored rows (cannot happen for a marker/face the view owns) stays memory-safe
  const int by = min(max((int)wy - g.yoff, -(GUARD_ROWS - 1)), g.ny + GUARD_ROWS - 3);

  const bool ok00 = face_has<TYPE>(fluid, g, bx, by);
  const bool ok10 = face_has<TYPE>(fluid, g, bx + 1, by);
  const bool ok01 = face_has<TYPE>(fluid, g, bx, by + 1);
  const bool ok11 = face_has<TYPE>(fluid, g, bx + 1, by + 1);
  const size_t c = gidx(g, bx, by);
  const float q00 = ok00 ? q[c] : 0.f;
  const float q10 = ok10 ? q[c + 1] : 0.f;
  const float q01 = ok01 ? q[c + g.pitch] : 0.f;
  const float q11 = ok11 ? q[c + g.pitch + 1] : 0.f;

  const float left = lerp_ref(q00, q01, snap_fraction(fy, ok00, ok01));
  const float right = lerp_ref(q10, q11, snap_fraction(fy, ok10, ok11));
  return lerp_ref(left, right, snap_fraction(fx, ok00 | ok01, ok10 | ok11));
}

}  // namespace euler
