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
  float p_x, p_y;       // P cells (the --rainbow colour planes)
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

// bilinear() of main.c:318-331 on four corner values with their usability flags: unusable
// corners read as 0 (main.c:333-335) and are excluded by snapping the fraction; vertical lerps
// first, then the horizontal one.
__device__ __forceinline__ float bilinear_masked(float v00, float v10, float v01, float v11,
                                                 bool ok00, bool ok10, bool ok01, bool ok11,
                                                 float fx, float fy) {
  const float q00 = ok00 ? v00 : 0.f;
  const float q10 = ok10 ? v10 : 0.f;
  const float q01 = ok01 ? v01 : 0.f;
  const float q11 = ok11 ? v11 : 0.f;
  const float left = lerp_ref(q00, q01, snap_fraction(fy, ok00, ok01));
  const float right = lerp_ref(q10, q11, snap_fraction(fy, ok10, ok11));
  return lerp_ref(left, right, snap_fraction(fx, ok00 | ok01, ok10 | ok11));
}

// One 64-bit index per sample; every other address is that pointer plus an immediate or plus
// the pitch.  The four values are loaded unconditionally (every address is inside the
// allocation: guard rows, clamped indices) and unusable corners are replaced by 0 afterwards, so
// the value loads do not wait for the mask loads — same result as reading 0 for an unusable
// corner (main.c:333-335), shorter dependency chain.  A U sample needs the fluid flags of a
// 3 x 2 block of cells, a V sample of a 2 x 3 block; each flag is loaded once.
template <int TYPE>
__device__ __forceinline__ float interpolate(const float* __restrict__ q,
                                             const uint8_t* __restrict__ fluid,
                                             const Grid& g, const InterpLimits& lim,
                                             float ix, float iy) {
  const float hx = TYPE == FACE_U ? lim.u_x : (TYPE == FACE_V ? lim.v_x : lim.p_x);
  const float hy = TYPE == FACE_U ? lim.u_y : (TYPE == FACE_V ? lim.v_y : lim.p_y);
  ix = ix < 0.f ? 0.f : (ix > hx ? hx : ix);
  iy = iy < 0.f ? 0.f : (iy > hy ? hy : iy);
  const float wx = truncf(ix), wy = truncf(iy);
  const float fx = ix - wx, fy = iy - wy;                    // == modff for finite input
  const int bx = (int)wx;
  // global row -> row of this view; clamped into the guard band so that a sample outside the
  // locally stored rows (cannot happen for a marker/face the view owns) stays memory-safe
  const int by = min(max((int)wy - g.yoff, -(GUARD_ROWS - 1)), g.ny + GUARD_ROWS - 3);

  const long c = (long)by * g.pitch + bx;
  const float* q0 = q + c;
  const float* q1 = q0 + g.pitch;
  const float v00 = q0[0], v10 = q0[1], v01 = q1[0], v11 = q1[1];
  const uint8_t* m0 = fluid + c;
  const uint8_t* m1 = m0 + g.pitch;
  bool ok00, ok10, ok01, ok11;
  if (TYPE == FACE_U) {        // a U face touches fluid if P(x) or P(x+1) is fluid
    const bool a0 = m0[0] != 0, a1 = m0[1] != 0, a2 = m0[2] != 0;
    const bool b0 = m1[0] != 0, b1 = m1[1] != 0, b2 = m1[2] != 0;
    ok00 = a0 | a1; ok10 = a1 | a2; ok01 = b0 | b1; ok11 = b1 | b2;
  } else if (TYPE == CELL_P) { // a P sample is usable iff its cell is fluid
    ok00 = m0[0] != 0; ok10 = m0[1] != 0; ok01 = m1[0] != 0; ok11 = m1[1] != 0;
  } else {                     // a V face if P(y) or P(y+1) is
    const uint8_t* m2 = m1 + g.pitch;
    const bool a0 = m0[0] != 0, a1 = m0[1] != 0, b0 = m1[0] != 0, b1 = m1[1] != 0;
    const bool c0 = m2[0] != 0, c1 = m2[1] != 0;
    ok00 = a0 | b0; ok10 = a1 | b1; ok01 = b0 | c0; ok11 = b1 | c1;
  }
  return bilinear_masked(v00, v10, v01, v11, ok00, ok10, ok01, ok11, fx, fy);
}

}  // namespace euler
