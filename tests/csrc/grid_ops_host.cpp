// tests/csrc/grid_ops_host.cpp — TEST INFRASTRUCTURE: the per-quad arithmetic of the MAC-grid
// stage kernels (euler_b200/csrc/grid_ops.cuh) compiled for the host and swept over padded host
// planes, for a bit-for-bit comparison with the oracle (tests/test_kernel_arith_host.py).  The wrappers
// below repeat only what the kernels add around a quad: zero-initialised results, the
// "neighbourhood holds fluid" test, the stores.  The fluid mask of a quad comes from
// load_quad_mask (plain loads) instead of the kernels' warp shuffle — same three words.
#include <math.h>
#include <stdint.h>

#include <algorithm>
using std::max;
using std::min;

#include "grid_ops.cuh"

using namespace euler;

namespace {

Grid make_grid(int nx, int ny, int pitch) {
  Grid g;
  g.nx = nx; g.ny = ny; g.pitch = pitch; g.yoff = 0; g.gny = ny; g.th = 32;
  return g;
}
InterpLimits make_limits(int nx, int ny) {
  InterpLimits lim;
  lim.u_x = nextafterf((float)(nx - 2), 0.f); lim.u_y = nextafterf((float)(ny - 1), 0.f);
  lim.v_x = nextafterf((float)(nx - 1), 0.f); lim.v_y = nextafterf((float)(ny - 2), 0.f);
  lim.p_x = nextafterf((float)(nx - 1), 0.f); lim.p_y = nextafterf((float)(ny - 1), 0.f);
  return lim;
}
F4 zero4() { F4 z; for (int k = 0; k < 4; ++k) z.v[k] = 0.f; return z; }

}  // namespace

extern "C" {

// (u, v) -> (uo, vo): extrapolate(u), extrapolate(v), zero_bounds(u), zero_bounds(v)  main.c:865-868
void ops_extrapolate_bounds(int nx, int ny, int pitch, const float* u, const float* v, const uint8_t* fluid,
                            const uint8_t* prev, const uint8_t* solid, float* uo, float* vo) {
  const Grid g = make_grid(nx, ny, pitch);
  for (int y = 0; y < ny; ++y)
    for (int x0 = 0; x0 < pitch; x0 += 4) {
      const size_t c = gidx(g, x0, y);
      const QuadMask f = load_quad_mask(fluid, g, c);
      F4 ru = zero4(), rv = zero4();
      if (f.any()) extrapolate_quad(g, f, x0, y, c, u, v, prev, solid, ru, rv);
      st_f4(uo + c, ru); st_f4(vo + c, rv);
    }
}

// (u, v) -> (uo, vo): advect_u, advect_v, apply_body_forces, zero_bounds(utmp), (vtmp)  main.c:871-889
void ops_advect_velocity(int nx, int ny, int pitch, const float* u, const float* v, const uint8_t* fluid,
                         const uint8_t* solid, float dt, float h, float gravity, float* uo, float* vo) {
  const Grid g = make_grid(nx, ny, pitch);
  const InterpLimits lim = make_limits(nx, ny);
  for (int y = 0; y < ny; ++y)
    for (int x0 = 0; x0 < pitch; x0 += 4) {
      const size_t c = gidx(g, x0, y);
      const QuadMask f = load_quad_mask(fluid, g, c);
      F4 ru = zero4(), rv = zero4();
      if (f.any()) advect_quad(g, lim, f, x0, y, c, u, v, fluid, solid, dt, h, gravity, ru, rv);
      st_f4(uo + c, ru); st_f4(vo + c, rv);
    }
}

// b, a_diag (fluid cells only), returns all_zero(b) == 0 ? 1 : 0  main.c:713-733, 742
int ops_build_rhs(int nx, int ny, int pitch, const float* u, const float* v, const uint8_t* fluid,
                  const uint8_t* solid, float h, double scale, double* b_out, int8_t* adiag) {
  const Grid g = make_grid(nx, ny, pitch);
  bool any = false;
  for (int y = 0; y < ny; ++y)
    for (int x0 = 0; x0 < pitch; x0 += 4) {
      const size_t c = gidx(g, x0, y);
      const unsigned mf = ld_u8x4(fluid + c);
      D4g b;
      for (int k = 0; k < 4; ++k) b.v[k] = 0.0;
      if (mf) any |= rhs_quad(g, mf, c, u, v, solid, adiag, h, scale, true, b);
      st_d4(b_out + c, b);
    }
  return any ? 1 : 0;
}

// clamp p >= 0 on fluid, (utmp, vtmp) - grad p -> (uo, vo)  main.c:769-805
void ops_pressure_update(int nx, int ny, int pitch, double* p, const float* ut, const float* vt,
                         const uint8_t* fluid, const uint8_t* solid, float dt, float kk, float* uo, float* vo) {
  const Grid g = make_grid(nx, ny, pitch);
  for (int y = 0; y < ny; ++y)
    for (int x0 = 0; x0 < pitch; x0 += 4) {
      const size_t c = gidx(g, x0, y);
      const QuadMask f = load_quad_mask(fluid, g, c);
      F4 ru = zero4(), rv = zero4();
      if (f.any()) pressure_quad(g, f, x0, y, c, p, ut, vt, solid, dt, kk, ru, rv);
      st_f4(uo + c, ru); st_f4(vo + c, rv);
    }
}

}  // extern "C"
