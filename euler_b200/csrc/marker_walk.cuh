// euler_b200/csrc/marker_walk.cuh — what the marker kernels do to ONE marker: the RK1 step with
// the reference's grid-line walk and solid rewind (advect_markers + velocity_at + time_to,
// main.c:440-537) and the marker -> cell map of refresh_marker_counts (main.c:106-107).  fp32,
// no FMA contraction, the reference's operation order: positions are bit-exact.  The same
// source compiles for the host, where tests/test_kernel_arith_host.py checks it against the oracle.
#pragma once
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "interp.cuh"

namespace euler {

namespace {

__device__ __forceinline__ float time_until(float from, float to, float vel) {
  return fabsf(vel) > 0.f ? (to - from) / vel : FLT_MAX;     // main.c:451-457
}

// A rewind (main.c:500-501 / 517-518) seen while walking one marker: it fires iff the
// crossing time t_hit is still < the marker's remaining dt, and then takes t_prev off dt.
struct HitRec {
  int n;
  float t_hit[2], t_prev[2];      // a component can be blocked only once -> at most 2 hits
};

// One marker, RK1 with the reference's grid-line walk (main.c:466-535).
template <bool RECORD>
__device__ __forceinline__ float2 walk_marker(const Grid& g, const InterpLimits& lim,
                                              const float* __restrict__ u,
                                              const float* __restrict__ v,
                                              const uint8_t* __restrict__ fluid,
                                              const uint8_t* __restrict__ solid, float h,
                                              float2 pos, float dt, HitRec* rec = nullptr) {
  if (RECORD) rec->n = 0;
  float px = pos.x, py = pos.y;
  // velocity_at, main.c:440-449
  float vx = interpolate<FACE_U>(u, fluid, g, lim, div_h(px, h) - 1.f, div_h(py, h) - 0.5f);
  float vy = interpolate<FACE_V>(v, fluid, g, lim, div_h(px, h) - 0.5f, div_h(py, h) - 1.f);

  int cx = (int)floorf(div_h(px, h));
  int cy = (int)floorf(div_h(py, h));
  const int step_x = vx > 0 ? 1 : -1;
  const int step_y = vy > 0 ? 1 : -1;
  int line_x = cx + (vx > 0 ? 1 : 0);
  int line_y = cy + (vy > 0 ? 1 : 0);
  const int off_x = vx < 0 ? -1 : 0;
  const int off_y = vy < 0 ? -1 : 0;
  float gx = line_x * h, gy = line_y * h;
  // (certifying "no grid line is crossed" without the two IEEE divisions of time_to — |line - p| >=
  // dt |v| (1 + 2^-20) implies fl((line - p) / v) > dt — and returning p + dt v at once was measured:
  // bit-identical, but advect_markers 2.80 -> 3.00 ms at 16384^2; the divisions are not what limits it)
  float tx = time_until(px, gx, vx);
  float ty = time_until(py, gy, vy);

  float t_prev = 0.f;
  float t_next = fminf(tx, ty);
  while (t_next < dt) {
    if (tx < ty) {
      const int sx = min(max(line_x + off_x, 0), g.nx - 1), sy = min(max(cy - g.yoff, 0), g.ny - 1);
      if (solid[gidx(g, sx, sy)]) {
        if (RECORD && rec->n < 2) { rec->t_hit[rec->n] = t_next; rec->t_prev[rec->n] = t_prev; rec->n++; }
        px = px + t_prev * vx; py = py + t_prev * vy;        // rewind, main.c:500
        dt -= t_prev;
        t_next = 0.f;
        vx = 0.f;
        tx = FLT_MAX;
        ty = time_until(py, gy, vy);
      } else {
        cx = line_x;
        line_x = cx + step_x;
        gx = line_x * h;
        tx = time_until(px, gx, vx);
      }
    } else {
      const int sx = min(max(cx, 0), g.nx - 1), sy = min(max(line_y + off_y - g.yoff, 0), g.ny - 1);
      if (solid[gidx(g, sx, sy)]) {
        if (RECORD && rec->n < 2) { rec->t_hit[rec->n] = t_next; rec->t_prev[rec->n] = t_prev; rec->n++; }
        px = px + t_prev * vx; py = py + t_prev * vy;        // main.c:517
        dt -= t_prev;
        t_next = 0.f;
        vy = 0.f;
        ty = FLT_MAX;
        tx = time_until(px, gx, vx);
      } else {
        cy = line_y;
        line_y = cy + step_y;
        gy = line_y * h;
        ty = time_until(py, gy, vy);
      }
    }
    t_prev = t_next;
    t_next = fminf(tx, ty);
  }
  const float t = (t_next < FLT_MAX) ? dt : t_prev;          // main.c:534
  return make_float2(px + t * vx, py + t * vy);
}

__device__ __forceinline__ bool marker_cell(const Grid& g, float h, float2 p, size_t* cell) {
  int cx = (int)floorf(div_h(p.x, h));                       // main.c:106-107
  int cy = (int)floorf(div_h(p.y, h));
  // the reference asserts 0 < x < X, 0 < y < Y (main.c:108, compiled out); clamp so a stray
  // marker lands in the sink ring and is deleted instead of indexing out of bounds
  cx = min(max(cx, 0), g.nx - 1);
  cy = min(max(cy - g.yoff, 0), g.ny - 1);                  // global row -> stored row
  *cell = gidx(g, cx, cy);
  return true;
}

}  // namespace

}  // namespace euler
