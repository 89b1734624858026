// tests/csrc/interp_host.cpp — TEST INFRASTRUCTURE: the masked bilinear sampler of the advection
// kernels (euler_b200/csrc/interp.cuh, reference main.c:301-364) compiled for the host and called
// on padded host planes, for a bit-for-bit comparison with the oracle's orc_interpolate
// (tests/test_kernel_arith_host.py).  Same source as the device code; g++ -ffp-contract=off.
#include <math.h>
#include <stdint.h>

#include <algorithm>
using std::max;
using std::min;

#include "interp.cuh"

using namespace euler;

extern "C" {

// out[i] = interpolate<type>(q, fluid, ix[i], iy[i]); planes in device layout (pitch, guard rows),
// pointers given for row 0.  type: 0 = P cell, 1 = U face, 2 = V face (celltype_t, main.c:46-50)
void ops_interpolate(int nx, int ny, int pitch, int type, const float* q, const uint8_t* fluid, int n,
                     const float* ix, const float* iy, float* out) {
  Grid g;
  g.nx = nx; g.ny = ny; g.pitch = pitch; g.yoff = 0; g.gny = ny; g.th = 32;
  InterpLimits lim;
  lim.u_x = nextafterf((float)(nx - 2), 0.f); lim.u_y = nextafterf((float)(ny - 1), 0.f);   // main.c:339-340
  lim.v_x = nextafterf((float)(nx - 1), 0.f); lim.v_y = nextafterf((float)(ny - 2), 0.f);
  lim.p_x = nextafterf((float)(nx - 1), 0.f); lim.p_y = nextafterf((float)(ny - 1), 0.f);
  for (int i = 0; i < n; ++i)
    out[i] = type == FACE_U ? interpolate<FACE_U>(q, fluid, g, lim, ix[i], iy[i])
           : type == FACE_V ? interpolate<FACE_V>(q, fluid, g, lim, ix[i], iy[i])
                            : interpolate<CELL_P>(q, fluid, g, lim, ix[i], iy[i]);
}

}  // extern "C"

// ---- the reference's random stream with jump-ahead (csrc/rng.cuh) -----------------------------
#include "rng.cuh"

extern "C" {

// state after k draws, by jump-ahead (GF(2) matrix powers)
unsigned long long ops_rng_jump(unsigned long long state, unsigned long long k) {
  static unsigned long long table[64 * 64];
  static bool built = false;
  if (!built) { rng_build_jump_table(table); built = true; }
  return rng_jump(table, state, k);
}
// n sequential draws as the kernels form them: step, then randf() of the new state; returns the state
unsigned long long ops_rng_draws(unsigned long long state, int n, float* out) {
  for (int i = 0; i < n; ++i) { state = rng_step(state); out[i] = rng_float(state); }
  return state;
}

}  // extern "C"

// ---- one marker: RK1 step with the grid-line walk, and the marker -> cell map -----------------
#include "marker_walk.cuh"

extern "C" {

// dst[i] = walk_marker(src[i], dt) for n markers (x, y pairs); every marker gets the full dt
// (EULER_MARKERS_FAST; the oracle's quirk_marker_dt_leak = 0).  cells[i] = the flat cell index
// refresh_marker_counts bins the MOVED marker into.
void ops_walk_markers(int nx, int ny, int pitch, const float* u, const float* v, const uint8_t* fluid,
                      const uint8_t* solid, float h, float dt, int n, const float* src, float* dst,
                      long long* cells) {
  Grid g;
  g.nx = nx; g.ny = ny; g.pitch = pitch; g.yoff = 0; g.gny = ny; g.th = 32;
  InterpLimits lim;
  lim.u_x = nextafterf((float)(nx - 2), 0.f); lim.u_y = nextafterf((float)(ny - 1), 0.f);
  lim.v_x = nextafterf((float)(nx - 1), 0.f); lim.v_y = nextafterf((float)(ny - 2), 0.f);
  lim.p_x = nextafterf((float)(nx - 1), 0.f); lim.p_y = nextafterf((float)(ny - 1), 0.f);
  for (int i = 0; i < n; ++i) {
    const float2 p = walk_marker<false>(g, lim, u, v, fluid, solid, h, make_float2(src[2 * i], src[2 * i + 1]), dt);
    dst[2 * i] = p.x; dst[2 * i + 1] = p.y;
    size_t c = 0;
    marker_cell(g, h, p, &c);
    cells[i] = (long long)(c / (size_t)pitch) * nx + (long long)(c % (size_t)pitch);   // back to [ny][nx] indexing
  }
}

}  // extern "C"
