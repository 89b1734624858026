/* oracle/euler_oracle.c — TEST INFRASTRUCTURE ONLY (see euler_oracle.h).
 *
 * CPU restatement of cgmb/euler's fluid solve with a run-time grid size.  Build with
 * -O2 -ffp-contract=off and WITHOUT -ffast-math so that every fp32/fp64 operation is a
 * single IEEE-754 round-to-nearest operation in source order — that is what the CUDA
 * kernels (compiled -fmad=false) reproduce bit for bit where the domain allows.
 *
 * "ref" below = /root/reference/main.c unless another file is named.
 */
#include "euler_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { CELL_P = 0, FACE_U = 1, FACE_V = 2 };   /* ref celltype_t :46-50 */

#define IDX(o, x, y) ((size_t)(y) * (size_t)(o)->nx + (size_t)(x))

/* ---------------------------------------------------------------- allocation ---- */

static void *zalloc(size_t n, size_t sz) {
  void *p = calloc(n ? n : 1, sz);
  if (!p) abort();
  return p;
}

orc_sim *orc_create(int nx, int ny) {
  orc_sim *o = zalloc(1, sizeof *o);
  size_t n = (size_t)nx * ny;
  o->nx = nx; o->ny = ny;
  o->h = 1.f; o->rho = 1.f; o->gravity = -10.f;          /* ref :58-60 */
  o->max_iterations = 100;                                /* ref :735 */
  o->tol = 1e-6f;                                         /* ref :736 (float literal) */
  o->precon_mode = ORC_PRECON_IC0;
  o->quirk_marker_dt_leak = 1;
  o->u = zalloc(n, 4); o->v = zalloc(n, 4); o->utmp = zalloc(n, 4); o->vtmp = zalloc(n, 4);
  o->solid = zalloc(n, 1); o->source = zalloc(n, 1); o->sink = zalloc(n, 1);
  o->count = zalloc(n, 1); o->prev_count = zalloc(n, 1);
  o->max_markers = 4 * n;                                 /* ref :92 */
  o->markers = zalloc(o->max_markers, sizeof(orc_vec2));
  o->rng_state = 0x9bd185c449534b91ull;                   /* ref :204 */
  o->adiag = zalloc(n, 1);
  o->precon = zalloc(n, 8); o->q = zalloc(n, 8);
  o->b = zalloc(n, 8); o->p = zalloc(n, 8); o->r = zalloc(n, 8);
  o->z = zalloc(n, 8); o->s = zalloc(n, 8);
  o->cr = zalloc(n, 4); o->cg = zalloc(n, 4); o->cb = zalloc(n, 4);
  o->crtmp = zalloc(n, 4); o->cgtmp = zalloc(n, 4); o->cbtmp = zalloc(n, 4);
  o->r32 = zalloc(n, 4); o->z32 = zalloc(n, 4); o->s32 = zalloc(n, 4);
  o->q32 = zalloc(n, 4); o->as32 = zalloc(n, 4); o->pc32 = zalloc(n, 4);
  return o;
}

void orc_destroy(orc_sim *o) {
  if (!o) return;
  free(o->u); free(o->v); free(o->utmp); free(o->vtmp);
  free(o->solid); free(o->source); free(o->sink); free(o->count); free(o->prev_count);
  free(o->markers); free(o->adiag); free(o->precon); free(o->q);
  free(o->b); free(o->p); free(o->r); free(o->z); free(o->s);
  free(o->cr); free(o->cg); free(o->cb); free(o->crtmp); free(o->cgtmp); free(o->cbtmp);
  free(o->r32); free(o->z32); free(o->s32); free(o->q32); free(o->as32); free(o->pc32);
  free(o);
}

/* ----------------------------------------------------------------------- RNG ---- */

/* xorshift64* keeping the high 32 bits: ref misc/rng.c:5-20 */
static uint32_t next_u32(uint64_t *state) {
  uint64_t s = *state;
  s ^= s >> 12;
  s ^= s << 25;
  s ^= s >> 27;
  *state = s;
  return (uint32_t)((s * 0x2545F4914F6CDD1Dull) >> 32);
}

/* ref :203-207 — the divide is done in double, the result can be exactly 1.0f */
float orc_randf(orc_sim *o) {
  uint32_t bits = next_u32(&o->rng_state);
  o->rng_draws++;
  return (float)(bits / (double)UINT32_MAX);
}

/* --------------------------------------------------------- cell predicates ---- */

/* ref :119-147: a P cell has the property itself; a U face has it if either of the two
 * P cells it separates (x, x+1) has it; a V face if either of (y, y+1) has it. */
static int has(const orc_sim *o, const uint8_t *mask, int x, int y, int type) {
  int a = mask[IDX(o, x, y)] != 0;
  if (type == FACE_U) return a | (mask[IDX(o, x + 1, y)] != 0);
  if (type == FACE_V) return a | (mask[IDX(o, x, y + 1)] != 0);
  return a;
}

/* ref :149-156 */
static void extent(const orc_sim *o, int type, int *sx, int *sy) {
  *sx = o->nx - (type == FACE_U);
  *sy = o->ny - (type == FACE_V);
}

/* ------------------------------------------------------------------ markers ---- */

/* ref :102-117.  prev <- cur, cur <- 0, re-bin; markers found in a sink or solid cell are
 * removed by moving the last marker into their slot (which is then examined next). */
void orc_refresh_marker_counts(orc_sim *o) {
  size_t n = (size_t)o->nx * o->ny;
  memcpy(o->prev_count, o->count, n);
  memset(o->count, 0, n);
  size_t i = 0;
  while (i < o->n_markers) {
    int cx = (int)floorf(o->markers[i].x / o->h);
    int cy = (int)floorf(o->markers[i].y / o->h);
    size_t c = IDX(o, cx, cy);
    if (o->sink[c] || o->solid[c]) {
      o->markers[i] = o->markers[--o->n_markers];
    } else {
      o->count[c]++;          /* uint8: wraps at 256, ref :96,114 */
      i++;
    }
  }
}

/* ref :209-274, minus colour */
void orc_init_from_text(orc_sim *o, const char *text, int length) {
  const int nx = o->nx, ny = o->ny;
  uint8_t *fluid = zalloc((size_t)nx * ny, 1);
  int pos = 0;
  /* first text row is y = ny-2, first column is x = 1 (ref :220-222) */
  for (int y = ny - 2; y > 0 && pos < length; --y) {
    int x = 1;
    while (x < nx - 1 && pos < length) {
      char c = text[pos++];
      if (c == '\n') break;
      size_t k = IDX(o, x, y);
      switch (c) {
        case 'X': o->solid[k] = 1; break;
        case '0': fluid[k] = 1; break;
        case '?': fluid[k] = 1; o->source[k] = 1; break;
        case '=': o->sink[k] = 1; break;
        default: break;
      }
      ++x;
    }
    if (x == nx - 1) {               /* over-long line: drop the rest (ref :238-240) */
      while (pos < length && text[pos++] != '\n') {}
    }
  }
  /* ring of sinks (ref :244-252) */
  for (int y = 0; y < ny; ++y) { o->sink[IDX(o, 0, y)] = 1; o->sink[IDX(o, nx - 1, y)] = 1; }
  for (int x = 0; x < nx; ++x) { o->sink[IDX(o, x, 0)] = 1; o->sink[IDX(o, x, ny - 1)] = 1; }
  /* 4 jittered markers per fluid cell, columns outermost, x drawn before y (ref :255-266) */
  size_t m = 0;
  for (int x = 0; x < nx; ++x) {
    for (int y = 0; y < ny; ++y) {
      if (!fluid[IDX(o, x, y)]) continue;
      for (int k = 0; k < 4; ++k) {
        float px = x + (k < 2 ? 0 : 0.5f) + (orc_randf(o) / 2);
        float py = y + (k % 2 ? 0 : 0.5f) + (orc_randf(o) / 2);
        o->markers[m].x = o->h * px;
        o->markers[m].y = o->h * py;
        m++;
      }
    }
  }
  o->n_markers = m;
  free(fluid);
  orc_refresh_marker_counts(o);
  if (o->rainbow) orc_colorize(o);                           /* ref :271-273 */
}

/* ------------------------------------------------------------------ rainbow ---- */

/* misc/color.h hsv_basis: periodic in t with period 6, values in [0,1] */
float orc_hsv_basis(float t) {
  t -= 6.f * floorf(1.f / 6 * t);
  if (t < 0.f) t += 6.f;
  if (t < 1.f) return t;
  if (t < 3.f) return 1.f;
  if (t < 4.f) return 4.f - t;
  return 0.f;
}

/* ref :187-201: hue ramps along x+y with a period of 60 cells; source cells start at t = 0 */
void orc_colorize(orc_sim *o) {
  for (int y = 0; y < o->ny; ++y) {
    for (int x = 0; x < o->nx; ++x) {
      size_t c = IDX(o, x, y);
      if (!o->count[c]) continue;
      float t = 0.f;
      if (!o->source[c]) t = (x + y) * 6.f / 60.f;           /* k_initial_color_period, ref :84 */
      o->cr[c] = orc_hsv_basis(t + 2.f);
      o->cg[c] = orc_hsv_basis(t);
      o->cb[c] = orc_hsv_basis(t - 2.f);
    }
  }
}

/* ref :424-438: back-trace from the cell centre with the mean of the two faces either side */
void orc_advect_p(const orc_sim *o, const float *q, const float *u, const float *v, float dt, float *out) {
  for (int y = 0; y < o->ny; ++y) {
    for (int x = 0; x < o->nx; ++x) {
      size_t c = IDX(o, x, y);
      if (!o->count[c]) continue;
      float dy = (v[c] + v[IDX(o, x, y - 1)]) / 2;
      float dx = (u[c] + u[IDX(o, x - 1, y)]) / 2;
      float px = x - dx * dt / o->h;
      float py = y - dy * dt / o->h;
      out[c] = orc_interpolate(o, q, px, py, CELL_P);
    }
  }
}

/* ref :276-298.  Argument evaluation order of
 * v2f(x+randf(), y+randf()) is unspecified in C; gcc 13 on x86-64 (the build that the
 * parity oracle oracle/_ref is made with) evaluates the SECOND argument first, i.e. the
 * y jitter takes the earlier draw.  tests/test_oracle.py pins this. */
void orc_update_fluid_sources(orc_sim *o) {
  const size_t cap = o->max_markers - 1;
  o->source_exhausted |= (o->n_markers == cap);
  const float t = 0.6f / 10.f * o->frame_count;              /* k_source_color_period, ref :83, :283 */
  for (int y = 0; y < o->ny; ++y) {
    for (int x = 0; x < o->nx; ++x) {
      size_t c = IDX(o, x, y);
      if (!o->source[c]) continue;
      if (!o->source_exhausted && o->count[c] < 4) {
        float jy = orc_randf(o);
        float jx = orc_randf(o);
        o->markers[o->n_markers].x = o->h * (x + jx);
        o->markers[o->n_markers].y = o->h * (y + jy);
        o->n_markers++;
        o->count[c]++;
        o->source_exhausted |= (o->n_markers == cap);
      }
      /* ref :292-294: written whether or not --rainbow is on (only read when it is) */
      o->cr[c] = orc_hsv_basis(t + 2.f);
      o->cg[c] = orc_hsv_basis(t);
      o->cb[c] = orc_hsv_basis(t - 2.f);
    }
  }
}

/* ------------------------------------------------ masked bilinear sampling ---- */

/* ref :311-313 */
static float lerp1(float a, float b, float f) { return (1.f - f) * a + f * b; }

/* ref :301-309: a missing end-point snaps the fraction onto the other one */
static float snap(float f, int lo_ok, int hi_ok) {
  if (!lo_ok) return 1.f;
  if (!hi_ok) return 0.f;
  return f;
}

/* ref :337-364 + :318-331.  Corner (i,j) = (base.x+i, base.y+j); a corner is usable iff it
 * is a fluid cell/face; unusable corners read as 0 and are excluded by snapping.  Vertical
 * lerps first, then the horizontal one. */
float orc_interpolate(const orc_sim *o, const float *q, float ix, float iy, int type) {
  int sx, sy; extent(o, type, &sx, &sy);
  float hi_x = nextafterf((float)(sx - 1), 0.f);
  float hi_y = nextafterf((float)(sy - 1), 0.f);
  ix = ix < 0.f ? 0.f : (ix > hi_x ? hi_x : ix);
  iy = iy < 0.f ? 0.f : (iy > hi_y ? hi_y : iy);
  float wx, wy;
  float fx = modff(ix, &wx);
  float fy = modff(iy, &wy);
  int bx = (int)wx, by = (int)wy;

  int ok00 = has(o, o->count, bx,     by,     type);
  int ok10 = has(o, o->count, bx + 1, by,     type);
  int ok01 = has(o, o->count, bx,     by + 1, type);
  int ok11 = has(o, o->count, bx + 1, by + 1, type);
  float q00 = ok00 ? q[IDX(o, bx,     by)]     : 0.f;
  float q10 = ok10 ? q[IDX(o, bx + 1, by)]     : 0.f;
  float q01 = ok01 ? q[IDX(o, bx,     by + 1)] : 0.f;
  float q11 = ok11 ? q[IDX(o, bx + 1, by + 1)] : 0.f;

  float left  = lerp1(q00, q01, snap(fy, ok00, ok01));
  float right = lerp1(q10, q11, snap(fy, ok10, ok11));
  return lerp1(left, right, snap(fx, ok00 | ok01, ok10 | ok11));
}

/* ------------------------------------------------- marker advection (RK1) ---- */

/* ref :451-457 */
static float time_until(float from, float to, float vel) {
  return fabsf(vel) > 0.f ? (to - from) / vel : FLT_MAX;
}

/* ref :464-537 with velocity_at :440-449.  Walks from grid line to grid line while the
 * next crossing happens before dt; on entering a solid cell the marker is rewound to the
 * previous crossing, the blocked velocity component is dropped and the walk restarts
 * with the remaining time. */
void orc_advect_markers(orc_sim *o, float dt_in) {
  const float h = o->h;
  /* QUIRK (ref :464, :501, :518): `dt` is the function parameter and `dt -= t_prev` is
   * never undone, so time consumed by one marker's rewind is also taken away from EVERY
   * LATER marker of the array in this sub-step.  This makes the array order observable.
   * quirk_marker_dt_leak=1 (default) reproduces it; 0 gives each marker the full dt. */
  float dt = dt_in;
  for (size_t i = 0; i < o->n_markers; ++i) {
    if (!o->quirk_marker_dt_leak) dt = dt_in;
    float px = o->markers[i].x, py = o->markers[i].y;
    float vx = orc_interpolate(o, o->u, px / h - 1.f,  py / h - 0.5f, FACE_U);
    float vy = orc_interpolate(o, o->v, px / h - 0.5f, py / h - 1.f,  FACE_V);

    int cx = (int)floorf(px / h);
    int cy = (int)floorf(py / h);
    int step_x = vx > 0 ? 1 : -1;
    int step_y = vy > 0 ? 1 : -1;
    int line_x = cx + (vx > 0 ? 1 : 0);         /* index of the next vertical grid line */
    int line_y = cy + (vy > 0 ? 1 : 0);
    int cell_off_x = vx < 0 ? -1 : 0;           /* cell entered when crossing line_x */
    int cell_off_y = vy < 0 ? -1 : 0;
    float gx = line_x * h, gy = line_y * h;
    float tx = time_until(px, gx, vx);
    float ty = time_until(py, gy, vy);

    float t_prev = 0.f;
    float t_next = fminf(tx, ty);
    while (t_next < dt) {
      if (tx < ty) {
        if (o->solid[IDX(o, line_x + cell_off_x, cy)]) {
          px = px + t_prev * vx; py = py + t_prev * vy;
          dt -= t_prev;
          t_next = 0;
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
        if (o->solid[IDX(o, cx, line_y + cell_off_y)]) {
          px = px + t_prev * vx; py = py + t_prev * vy;
          dt -= t_prev;
          t_next = 0;
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
    float t = (t_next < FLT_MAX) ? dt : t_prev;
    o->markers[i].x = px + t * vx;
    o->markers[i].y = py + t * vy;
  }
}

/* ------------------------------------------- extrapolation / boundaries ---- */

/* ref :173-185 with :158-171.  A face that touches fluid now but did not last step takes
 * the mean of those faces in its clamped 3x3 block that did.  In place; the faces written
 * are exactly those that no other face reads (they fail the prev test). */
void orc_extrapolate(orc_sim *o, float *q, int type) {
  int sx, sy; extent(o, type, &sx, &sy);
  for (int y = 0; y < sy; ++y) {
    for (int x = 0; x < sx; ++x) {
      if (has(o, o->prev_count, x, y, type) || !has(o, o->count, x, y, type)) continue;
      int x0 = x - 1 < 0 ? 0 : x - 1, x1 = x + 1 > sx - 1 ? sx - 1 : x + 1;
      int y0 = y - 1 < 0 ? 0 : y - 1, y1 = y + 1 > sy - 1 ? sy - 1 : y + 1;
      float sum = 0.f; int cnt = 0;
      for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx)
          if (has(o, o->prev_count, xx, yy, type)) { sum += q[IDX(o, xx, yy)]; cnt++; }
      q[IDX(o, x, y)] = sum / cnt;          /* cnt==0 -> NaN, as in the reference (assert off) */
    }
  }
}

/* ref :822-832 */
void orc_zero_bounds(const orc_sim *o, float *q, int type) {
  int sx, sy; extent(o, type, &sx, &sy);
  for (int y = 0; y < sy; ++y)
    for (int x = 0; x < sx; ++x)
      if (!has(o, o->count, x, y, type) || has(o, o->solid, x, y, type)) q[IDX(o, x, y)] = 0.f;
}

/* ------------------------------------------------ semi-Lagrangian advection ---- */

/* ref :382-399 (+ :378-380) */
void orc_advect_u(const orc_sim *o, const float *u, const float *v, float dt, float *out) {
  for (int y = 0; y < o->ny; ++y) {
    for (int x = 0; x < o->nx - 1; ++x) {
      if (!has(o, o->count, x, y, FACE_U)) continue;
      float dx = u[IDX(o, x, y)];
      float dy = orc_interpolate(o, v, x + 0.5f, y - 0.5f, FACE_V);
      float bx = x - dx * dt / o->h;
      float by = y - dy * dt / o->h;
      out[IDX(o, x, y)] = orc_interpolate(o, u, bx, by, FACE_U);
    }
  }
}

/* ref :405-422 (+ :401-403) */
void orc_advect_v(const orc_sim *o, const float *u, const float *v, float dt, float *out) {
  for (int y = 0; y < o->ny - 1; ++y) {
    for (int x = 0; x < o->nx; ++x) {
      if (!has(o, o->count, x, y, FACE_V)) continue;
      float dy = v[IDX(o, x, y)];
      float dx = orc_interpolate(o, u, x - 0.5f, y + 0.5f, FACE_U);
      float bx = x - dx * dt / o->h;
      float by = y - dy * dt / o->h;
      out[IDX(o, x, y)] = orc_interpolate(o, v, bx, by, FACE_V);
    }
  }
}

/* ref :539-545 */
void orc_apply_body_forces(const orc_sim *o, float *v, float dt) {
  for (int y = 0; y < o->ny - 1; ++y)
    for (int x = 0; x < o->nx; ++x)
      v[IDX(o, x, y)] += o->gravity * dt;
}

/* ref :808-820, :834-841 */
static float max_square(const orc_sim *o, const float *q, int type) {
  int sx, sy; extent(o, type, &sx, &sy);
  float best = 0;
  for (int y = 0; y < sy; ++y)
    for (int x = 0; x < sx; ++x) {
      float sq = q[IDX(o, x, y)] * q[IDX(o, x, y)];
      if (sq > best) best = sq;
    }
  return best;
}

float orc_calculate_timestep(const orc_sim *o, float frame_time) {
  const float reach = 0.75f * o->h;
  float vmax = sqrtf(max_square(o, o->u, FACE_U) + max_square(o, o->v, FACE_V));
  return fminf(reach / vmax, frame_time);
}

/* ------------------------------------------------------ pressure projection ---- */

#define FLUID(o, x, y) ((o)->count[IDX(o, x, y)] != 0)

/* ref :713-733.  Divergence is formed left to right in fp32 and only then widened. */
void orc_build_rhs(orc_sim *o, float dt, const float *u, const float *v) {
  const double scale = (o->h * o->h) * o->rho / dt;      /* fp32 expression, then widened */
  size_t n = (size_t)o->nx * o->ny;
  memset(o->b, 0, n * sizeof(double));
  for (int y = 0; y < o->ny; ++y) {
    for (int x = 0; x < o->nx; ++x) {
      if (!FLUID(o, x, y)) continue;
      size_t c = IDX(o, x, y);
      double div = (u[c] - u[c - 1] + v[c] - v[c - o->nx]) / o->h;
      o->b[c] = -div * scale;
      /* ref :554-559, :727-733: only fluid cells are (re)written */
      o->adiag[c] = (int8_t)(4 - o->solid[c - 1] - o->solid[c + 1]
                               - o->solid[c - o->nx] - o->solid[c + o->nx]);
    }
  }
}

/* ref :580-627, natural (row-major) ordering.  NB: inside the fluid branch
 * the "minus" couplings are always -1, so the diagonal recurrence reads precon at the left
 * and lower neighbours even when those are not fluid — values left over from earlier time
 * steps (precon is persistent and only ever written at fluid cells). */
static void precon_ic0(orc_sim *o, const double *r, double *z) {
  const int nx = o->nx, ny = o->ny;
  size_t n = (size_t)nx * ny;
  for (int y = 0; y < ny; ++y) {
    for (int x = 0; x < nx; ++x) {
      if (!FLUID(o, x, y)) continue;
      size_t c = IDX(o, x, y);
      double a = o->adiag[c];
      double wl = -1 * o->precon[c - 1];
      double wd = -1 * o->precon[c - nx];
      double e = a - wl * wl - wd * wd;
      if (e < 0.25 * a) e = a != 0 ? a : 1;
      o->precon[c] = 1 / sqrt(e);
    }
  }
  memset(o->q, 0, n * sizeof(double));
  for (int y = 0; y < ny; ++y) {
    for (int x = 0; x < nx; ++x) {
      if (!FLUID(o, x, y)) continue;
      size_t c = IDX(o, x, y);
      double t = r[c] - (-1 * o->precon[c - 1]) * o->q[c - 1]
                      - (-1 * o->precon[c - nx]) * o->q[c - nx];
      o->q[c] = t * o->precon[c];
    }
  }
  memset(z, 0, n * sizeof(double));
  for (int y = ny; y--;) {
    for (int x = nx; x--;) {
      if (!FLUID(o, x, y)) continue;
      size_t c = IDX(o, x, y);
      int ar = FLUID(o, x + 1, y) ? -1 : 0;
      int au = FLUID(o, x, y + 1) ? -1 : 0;
      double t = o->q[c] - ar * o->precon[c] * z[c + 1]
                         - au * o->precon[c] * z[c + nx];
      z[c] = t * o->precon[c];
    }
  }
}

/* NOT in the reference.  IC(0) of the same matrix under red-black ordering (red:
 * (x+y) even, eliminated first).  Red rows have no earlier neighbour, so E_red = a; black
 * rows see only red neighbours.  Same sigma=0.25 safety rule as ref :594-596.  Every
 * neighbour sum is taken in the fixed order left, right, down, up so the CUDA kernel can
 * reproduce it bit for bit.  Unlike the natural-order factor there are no stale reads:
 * only fluid neighbours contribute. */
static double rb_e_red(const orc_sim *o, size_t c) {
  double a = o->adiag[c];
  return a != 0 ? a : 1;
}
static void precon_redblack(orc_sim *o, const double *r, double *z) {
  const int nx = o->nx, ny = o->ny;
  size_t n = (size_t)nx * ny;
  const long off[4] = { -1, 1, -(long)nx, (long)nx };
  /* diagonal */
  for (int y = 1; y < ny - 1; ++y)
    for (int x = 1; x < nx - 1; ++x) {
      size_t c = IDX(o, x, y);
      if (!o->count[c]) continue;
      if (((x + y) & 1) == 0) { o->precon[c] = 1 / sqrt(rb_e_red(o, c)); continue; }
      double a = o->adiag[c];
      double e = a;
      for (int k = 0; k < 4; ++k) {
        size_t nb = c + off[k];
        if (o->count[nb]) e = e - 1 / rb_e_red(o, nb);
      }
      if (e < 0.25 * a) e = a != 0 ? a : 1;
      o->precon[c] = 1 / sqrt(e);
    }
  memset(o->q, 0, n * sizeof(double));
  memset(z, 0, n * sizeof(double));
  /* L q = r : red, then black */
  for (int colour = 0; colour < 2; ++colour)
    for (int y = 1; y < ny - 1; ++y)
      for (int x = 1; x < nx - 1; ++x) {
        size_t c = IDX(o, x, y);
        if (!o->count[c] || ((x + y) & 1) != colour) continue;
        double t = r[c];
        if (colour == 1)
          for (int k = 0; k < 4; ++k) {
            size_t nb = c + off[k];
            if (o->count[nb]) t = t + o->precon[nb] * o->q[nb];
          }
        o->q[c] = t * o->precon[c];
      }
  /* L^T z = q : black, then red */
  for (int colour = 1; colour >= 0; --colour)
    for (int y = 1; y < ny - 1; ++y)
      for (int x = 1; x < nx - 1; ++x) {
        size_t c = IDX(o, x, y);
        if (!o->count[c] || ((x + y) & 1) != colour) continue;
        double t = o->q[c];
        if (colour == 0)
          for (int k = 0; k < 4; ++k) {
            size_t nb = c + off[k];
            if (o->count[nb]) t = t + o->precon[c] * z[nb];
          }
        z[c] = t * o->precon[c];
      }
}

void orc_apply_preconditioner(orc_sim *o, const double *r, double *z) {
  if (o->precon_mode == ORC_PRECON_REDBLACK) precon_redblack(o, r, z);
  else precon_ic0(o, r, z);
}

/* ref :679-691 */
void orc_apply_a(const orc_sim *o, const double *s, double *out) {
  for (int y = 0; y < o->ny; ++y)
    for (int x = 0; x < o->nx; ++x) {
      if (!FLUID(o, x, y)) continue;
      size_t c = IDX(o, x, y);
      out[c] = o->adiag[c] * s[c]
             - (FLUID(o, x + 1, y) ? s[c + 1] : 0)
             - (FLUID(o, x, y + 1) ? s[c + o->nx] : 0)
             - (FLUID(o, x - 1, y) ? s[c - 1] : 0)
             - (FLUID(o, x, y - 1) ? s[c - o->nx] : 0);
    }
}

/* ref :629-639 — sequential row-major sum */
double orc_dot(const orc_sim *o, const double *a, const double *b) {
  double total = 0.f;
  size_t n = (size_t)o->nx * o->ny;
  for (size_t c = 0; c < n; ++c)
    if (o->count[c]) total += a[c] * b[c];
  return total;
}

/* ref :654-667 */
double orc_inf_norm(const orc_sim *o, const double *r) {
  double best = 0.f;
  size_t n = (size_t)o->nx * o->ny;
  for (size_t c = 0; c < n; ++c)
    if (o->count[c]) { double a = fabs(r[c]); if (a > best) best = a; }
  return best;
}

/* ref :641-652 */
int orc_all_zero(const orc_sim *o, const double *r) {
  size_t n = (size_t)o->nx * o->ny;
  for (size_t c = 0; c < n; ++c)
    if (o->count[c] && r[c] != 0.f) return 0;
  return 1;
}

/* ref :769-805 (+ accel :705-707): clamp p >= 0 on fluid, then subtract the pressure
 * gradient on faces that touch fluid and no solid; every other face is zeroed.  The
 * pressure difference is taken in fp64 and narrowed to fp32 as accel()'s argument. */
void orc_pressure_update(orc_sim *o, float dt, const float *u, const float *v, float *uout, float *vout) {
  size_t n = (size_t)o->nx * o->ny;
  for (size_t c = 0; c < n; ++c)
    if (o->count[c] && o->p[c] < 0.f) o->p[c] = 0.f;
  const float k = 1.f / (o->rho * o->h);
  for (int y = 0; y < o->ny; ++y)
    for (int x = 0; x < o->nx - 1; ++x) {
      size_t c = IDX(o, x, y);
      if (has(o, o->solid, x, y, FACE_U)) uout[c] = 0.f;
      else if (has(o, o->count, x, y, FACE_U)) {
        float dp = (float)(o->p[c + 1] - o->p[c]);
        uout[c] = u[c] + (-k * dp) * dt;
      } else uout[c] = 0.f;
    }
  for (int y = 0; y < o->ny - 1; ++y)
    for (int x = 0; x < o->nx; ++x) {
      size_t c = IDX(o, x, y);
      if (has(o, o->solid, x, y, FACE_V)) vout[c] = 0.f;
      else if (has(o, o->count, x, y, FACE_V)) {
        float dp = (float)(o->p[c + o->nx] - o->p[c]);
        vout[c] = v[c] + (-k * dp) * dt;
      } else vout[c] = 0.f;
    }
}

/* ------------------------------------------- mixed-precision PCG (NOT in the reference) ----
 * CPU mirror of the GPU's pcg_dtype = FP32 mode (red-black preconditioner
 * only).  Same algorithm and the same order of operations as project()'s loop (ref :735-767)
 * and precon_redblack above, but r, z, s, q, A s and the preconditioner diagonal are fp32
 * planes and every element-wise operation is ONE fp32 IEEE operation in source order (the
 * CUDA kernels reproduce them bit for bit).  What stays fp64: the pressure p (it accumulates
 * ~100 updates), the products and sums of the dot products, and sigma / alpha / beta (each
 * narrowed to fp32 once, where it multiplies a plane). */

/* the fp64 factor of precon_redblack, narrowed once */
void orc_rb_build32(orc_sim *o) {
  const int nx = o->nx, ny = o->ny;
  const long off[4] = { -1, 1, -(long)nx, (long)nx };
  for (int y = 1; y < ny - 1; ++y)
    for (int x = 1; x < nx - 1; ++x) {
      size_t c = IDX(o, x, y);
      if (!o->count[c]) continue;
      if (((x + y) & 1) == 0) { o->pc32[c] = (float)(1 / sqrt(rb_e_red(o, c))); continue; }
      double a = o->adiag[c];
      double e = a;
      for (int k = 0; k < 4; ++k) {
        size_t nb = c + off[k];
        if (o->count[nb]) e = e - 1 / rb_e_red(o, nb);
      }
      if (e < 0.25 * a) e = a != 0 ? a : 1;
      o->pc32[c] = (float)(1 / sqrt(e));
    }
}

void orc_rb_apply32(orc_sim *o, const float *r, float *z) {
  const int nx = o->nx, ny = o->ny;
  size_t n = (size_t)nx * ny;
  const long off[4] = { -1, 1, -(long)nx, (long)nx };
  const float *pc = o->pc32;
  float *q = o->q32;
  memset(q, 0, n * sizeof(float));
  memset(z, 0, n * sizeof(float));
  for (int colour = 0; colour < 2; ++colour)
    for (int y = 1; y < ny - 1; ++y)
      for (int x = 1; x < nx - 1; ++x) {
        size_t c = IDX(o, x, y);
        if (!o->count[c] || ((x + y) & 1) != colour) continue;
        float t = r[c];
        if (colour == 1)
          for (int k = 0; k < 4; ++k) {
            size_t nb = c + off[k];
            if (o->count[nb]) t = t + pc[nb] * q[nb];
          }
        q[c] = t * pc[c];
      }
  for (int colour = 1; colour >= 0; --colour)
    for (int y = 1; y < ny - 1; ++y)
      for (int x = 1; x < nx - 1; ++x) {
        size_t c = IDX(o, x, y);
        if (!o->count[c] || ((x + y) & 1) != colour) continue;
        float t = q[c];
        if (colour == 0)
          for (int k = 0; k < 4; ++k) {
            size_t nb = c + off[k];
            if (o->count[nb]) t = t + pc[c] * z[nb];
          }
        z[c] = t * pc[c];
      }
}

void orc_apply_a32(const orc_sim *o, const float *s, float *out) {
  for (int y = 0; y < o->ny; ++y)
    for (int x = 0; x < o->nx; ++x) {
      if (!FLUID(o, x, y)) continue;
      size_t c = IDX(o, x, y);
      float t = (float)o->adiag[c] * s[c];
      t = t - (FLUID(o, x + 1, y) ? s[c + 1] : 0.f);
      t = t - (FLUID(o, x, y + 1) ? s[c + o->nx] : 0.f);
      t = t - (FLUID(o, x - 1, y) ? s[c - 1] : 0.f);
      t = t - (FLUID(o, x, y - 1) ? s[c + 0 - o->nx] : 0.f);
      out[c] = t;
    }
}

static double dot32(const orc_sim *o, const float *a, const float *b) {
  double total = 0.0;
  size_t n = (size_t)o->nx * o->ny;
  for (size_t c = 0; c < n; ++c)
    if (o->count[c]) total += (double)a[c] * (double)b[c];
  return total;
}

/* r32 <- b - A p evaluated in fp64 on the current p (residual replacement) */
static void true_residual32(orc_sim *o) {
  orc_apply_a(o, o->p, o->z);                /* z (fp64 plane) is free in this mode */
  size_t n = (size_t)o->nx * o->ny;
  for (size_t c = 0; c < n; ++c)
    if (o->count[c]) o->r32[c] = (float)(o->b[c] - o->z[c]);
}

static void pcg_mixed(orc_sim *o) {
  size_t n = (size_t)o->nx * o->ny;
  for (size_t c = 0; c < n; ++c) o->r32[c] = (float)o->b[c];
  orc_rb_build32(o);
  orc_rb_apply32(o, o->r32, o->z32);
  memcpy(o->s32, o->z32, n * sizeof(float));
  double sigma = dot32(o, o->z32, o->r32);
  for (int it = 0; it < o->max_iterations; ++it) {
    orc_apply_a32(o, o->s32, o->as32);
    double alpha = sigma / dot32(o, o->as32, o->s32);
    const float na = (float)-alpha;
    float best = 0.f;
    for (size_t c = 0; c < n; ++c) {
      if (!o->count[c]) continue;
      o->p[c] = o->p[c] + (double)o->s32[c] * alpha;
      o->r32[c] = o->r32[c] + o->as32[c] * na;
      float a = fabsf(o->r32[c]);
      if (a > best) best = a;
    }
    o->last_iterations = it + 1;
    o->last_residual = best;
    if (o->last_residual <= o->tol) break;
    if (o->refresh_every > 0 && (it + 1) % o->refresh_every == 0) true_residual32(o);
    orc_rb_apply32(o, o->r32, o->z32);
    double sigma_new = dot32(o, o->z32, o->r32);
    const float beta = (float)(sigma_new / sigma);
    for (size_t c = 0; c < n; ++c) if (o->count[c]) o->s32[c] = o->z32[c] + beta * o->s32[c];
    sigma = sigma_new;
  }
}

/* ref :709-806 */
void orc_project(orc_sim *o, float dt, const float *u, const float *v, float *uout, float *vout) {
  size_t n = (size_t)o->nx * o->ny;
  orc_build_rhs(o, dt, u, v);
  memset(o->p, 0, n * sizeof(double));
  memcpy(o->r, o->b, n * sizeof(double));
  o->last_iterations = 0;
  o->last_solve_skipped = orc_all_zero(o, o->r);
  if (!o->last_solve_skipped) {
    o->total_solves++;
    if (o->pcg_dtype == ORC_PCG_FP32) {
      pcg_mixed(o);
      o->total_iterations += o->last_iterations;
      orc_pressure_update(o, dt, u, v, uout, vout);
      return;
    }
    orc_apply_preconditioner(o, o->r, o->z);
    memcpy(o->s, o->z, n * sizeof(double));
    double sigma = orc_dot(o, o->z, o->r);
    for (int it = 0; it < o->max_iterations; ++it) {
      orc_apply_a(o, o->s, o->z);
      double alpha = sigma / orc_dot(o, o->z, o->s);
      for (size_t c = 0; c < n; ++c) if (o->count[c]) o->p[c] += o->s[c] * alpha;     /* ref :694-702, :753 */
      for (size_t c = 0; c < n; ++c) if (o->count[c]) o->r[c] += o->z[c] * -alpha;    /* ref :754 */
      o->last_iterations = it + 1;
      o->last_residual = orc_inf_norm(o, o->r);
      if (o->last_residual <= o->tol) break;
      orc_apply_preconditioner(o, o->r, o->z);
      double sigma_new = orc_dot(o, o->z, o->r);
      double beta = sigma_new / sigma;
      for (size_t c = 0; c < n; ++c) if (o->count[c]) o->s[c] = o->z[c] + beta * o->s[c];  /* ref :669-677 */
      sigma = sigma_new;
    }
    o->total_iterations += o->last_iterations;
  }
  orc_pressure_update(o, dt, u, v, uout, vout);
}

/* ------------------------------------------------------------- step driver ---- */

/* body of the sub-step loop, ref :855-893 */
void orc_substep(orc_sim *o, float dt) {
  const size_t plane = (size_t)o->nx * o->ny * sizeof(float);
  o->last_dt = dt;
  orc_advect_markers(o, dt);
  orc_refresh_marker_counts(o);
  if (o->rainbow) {                                          /* ref :859-863 */
    orc_extrapolate(o, o->cr, CELL_P);
    orc_extrapolate(o, o->cg, CELL_P);
    orc_extrapolate(o, o->cb, CELL_P);
  }
  orc_update_fluid_sources(o);
  orc_extrapolate(o, o->u, FACE_U);
  orc_extrapolate(o, o->v, FACE_V);
  orc_zero_bounds(o, o->u, FACE_U);
  orc_zero_bounds(o, o->v, FACE_V);
  orc_advect_u(o, o->u, o->v, dt, o->utmp);
  orc_advect_v(o, o->u, o->v, dt, o->vtmp);
  if (o->rainbow) {                                          /* ref :873-882: whole-plane copies */
    orc_advect_p(o, o->cr, o->u, o->v, dt, o->crtmp); memcpy(o->cr, o->crtmp, plane);
    orc_advect_p(o, o->cg, o->u, o->v, dt, o->cgtmp); memcpy(o->cg, o->cgtmp, plane);
    orc_advect_p(o, o->cb, o->u, o->v, dt, o->cbtmp); memcpy(o->cb, o->cbtmp, plane);
  }
  orc_apply_body_forces(o, o->vtmp, dt);
  orc_zero_bounds(o, o->utmp, FACE_U);
  orc_zero_bounds(o, o->vtmp, FACE_V);
  orc_project(o, dt, o->utmp, o->vtmp, o->u, o->v);
  o->total_substeps++;
}

/* ref :843-900: at most 8 adaptive sub-steps per 0.1 s frame */
int orc_step_frame(orc_sim *o) {
  float frame_time = 0.1f;
  int steps = 0;
  for (; frame_time > 0.f && steps < 8; ++steps) {
    float dt = orc_calculate_timestep(o, frame_time);
    frame_time -= dt;
    orc_substep(o, dt);
  }
  o->frame_count++;                                          /* ref :899 */
  return steps;
}

uint64_t orc_fnv1a(const uint8_t *data, size_t n) {
  uint64_t hash = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { hash ^= data[i]; hash *= 1099511628211ull; }
  return hash;
}
