/* oracle/euler_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement, with a RUN-TIME grid size, of the per-timestep fluid solve of
 * cgmb/euler (reference main.c:102-900).  It exists to check the CUDA path; it is never
 * linked into, imported by, or called from the product (euler_b200/, libeuler_gpu.so).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Parity status: PINNED against the reference itself.  The reference has no tests or
 * golden vectors, but it compiles here unmodified (oracle/build_ref.sh →
 * oracle/_ref/ libraries); tests/test_oracle.py checks that every plane and the marker
 * array of this restatement are bit-identical to the reference's after whole frames and
 * after each individual stage, on all five shipped scenarios and on resampled ones.
 *
 * Every function cites the reference lines it follows.  Arrays are flat, row-major,
 * index = y*nx + x (the reference's [Y][X], main.c:64).
 */
#ifndef EULER_ORACLE_H
#define EULER_ORACLE_H
#include <stddef.h>
#include <stdint.h>

typedef struct orc_vec2 { float x, y; } orc_vec2;

enum { ORC_PRECON_IC0 = 0,      /* reference-faithful natural-order IC(0), main.c:580-627 */
       ORC_PRECON_REDBLACK = 1  /* NOT in the reference: red-black ordered IC(0); CPU
                                   mirror of the GPU-parallel mode, same arithmetic order */ };

enum { ORC_PCG_FP64 = 0, ORC_PCG_FP32 = 1 };

typedef struct orc_sim {
  int nx, ny;
  /* constants, main.c:58-60, 735-736, 838, 849-851 */
  float h, rho, gravity;
  int   max_iterations;
  double tol;
  int   precon_mode;
  int   quirk_marker_dt_leak;  /* 1 = reference behaviour (main.c:464,501,518), see .c */
  /* velocities, main.c:64-67 */
  float *u, *v, *utmp, *vtmp;
  /* static masks main.c:71-73, dynamic counts main.c:96-97 */
  uint8_t *solid, *source, *sink, *count, *prev_count;
  /* markers main.c:92-95 */
  orc_vec2 *markers;
  size_t n_markers, max_markers;
  int source_exhausted;
  /* RNG main.c:204 */
  uint64_t rng_state;
  uint64_t rng_draws;
  /* pressure solve: persistent planes main.c:552,577-578 + the stack VLAs of project() */
  int8_t *adiag;
  double *precon, *q;
  double *b, *p, *r, *z, *s;
  /* --rainbow colour transport (main.c:76-84, 89): a passive RGB scalar on the P cells */
  int    rainbow;              /* g_rainbow_enabled */
  float *cr, *cg, *cb, *crtmp, *cgtmp, *cbtmp;   /* g_r g_g g_b g_rtmp g_gtmp g_btmp */
  uint16_t frame_count;        /* g_frame_count (uint16, wraps), main.c:89 */
  /* bookkeeping for tests / benchmarks */
  int    last_iterations;      /* PCG iterations of the last project() (0 if skipped) */
  int    last_solve_skipped;   /* all_zero(r) fired, main.c:742 */
  double last_residual;        /* ||r||inf at exit */
  long   total_iterations, total_substeps, total_solves;
  float  last_dt;
  /* NOT in the reference: mixed-precision PCG, CPU mirror of the GPU's
   * euler_params.pcg_dtype = EULER_PCG_FP32 — red-black mode only.  The vectors r, z, s, q,
   * A s and the preconditioner diagonal are STORED and combined in fp32; the pressure p, the
   * dot products and the scalars alpha, beta, sigma stay fp64.  Every `refresh_every`
   * iterations (0 = never) r is replaced by the true residual b - A p evaluated in fp64. */
  int    pcg_dtype;            /* ORC_PCG_FP64 (default) / ORC_PCG_FP32 */
  int    refresh_every;
  float *r32, *z32, *s32, *q32, *as32, *pc32;
} orc_sim;

orc_sim *orc_create(int nx, int ny);
void     orc_destroy(orc_sim *o);
/* sim_init, main.c:209-274 (colorize() at the end when o->rainbow is set before the call) */
void     orc_init_from_text(orc_sim *o, const char *text, int length);
/* --rainbow pieces */
float    orc_hsv_basis(float t);                                          /* misc/color.h */
void     orc_colorize(orc_sim *o);                                        /* main.c:187-201 */
void     orc_advect_p(const orc_sim *o, const float *q, const float *u, const float *v, float dt, float *out); /* :424-438 */

/* stages, in sim_step order (main.c:851-893) */
float  orc_calculate_timestep(const orc_sim *o, float frame_time);       /* :834-841 */
void   orc_advect_markers(orc_sim *o, float dt);                          /* :464-537 */
void   orc_refresh_marker_counts(orc_sim *o);                             /* :102-117 */
void   orc_update_fluid_sources(orc_sim *o);                              /* :276-298 */
void   orc_extrapolate(orc_sim *o, float *q, int type);                   /* :173-185 */
void   orc_zero_bounds(const orc_sim *o, float *q, int type);             /* :822-832 */
void   orc_advect_u(const orc_sim *o, const float *u, const float *v, float dt, float *out); /* :382-399 */
void   orc_advect_v(const orc_sim *o, const float *u, const float *v, float dt, float *out); /* :405-422 */
void   orc_apply_body_forces(const orc_sim *o, float *v, float dt);       /* :539-545 */
void   orc_project(orc_sim *o, float dt, const float *u, const float *v, float *uout, float *vout); /* :709-806 */
void   orc_substep(orc_sim *o, float dt);                                 /* body of the loop :855-893 */
int    orc_step_frame(orc_sim *o);                                        /* sim_step :843-900; returns #sub-steps */

/* pieces of project(), exposed for per-kernel parity */
void   orc_build_rhs(orc_sim *o, float dt, const float *u, const float *v);   /* :713-733 */
void   orc_apply_preconditioner(orc_sim *o, const double *r, double *z);       /* :580-627 */
void   orc_apply_a(const orc_sim *o, const double *s, double *out);            /* :679-691 */
double orc_dot(const orc_sim *o, const double *a, const double *b);            /* :629-639 */
double orc_inf_norm(const orc_sim *o, const double *r);                        /* :654-667 */
int    orc_all_zero(const orc_sim *o, const double *r);                        /* :641-652 */
void   orc_pressure_update(orc_sim *o, float dt, const float *u, const float *v, float *uout, float *vout); /* :769-805 */
float  orc_interpolate(const orc_sim *o, const float *q, float ix, float iy, int type);  /* :337-364 */
float  orc_randf(orc_sim *o);                                                  /* :203-207 */

/* mixed-precision mirror pieces (fp32 planes), exposed for per-kernel parity */
void   orc_rb_build32(orc_sim *o);                                              /* -> pc32 */
void   orc_rb_apply32(orc_sim *o, const float *r, float *z);                    /* z = M^-1 r, scratch q32 */
void   orc_apply_a32(const orc_sim *o, const float *s, float *out);

/* 64-bit FNV-1a over a byte plane (BASELINE.md §3 known-answer hashes) */
uint64_t orc_fnv1a(const uint8_t *data, size_t n);
#endif
