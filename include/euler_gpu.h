/* include/euler_gpu.h — C-ABI of libeuler_gpu.so
 *
 * Drop-in boundary for ONE path of cgmb/euler: the per-timestep fluid solve, i.e.
 * `sim_step()` (reference main.c:843-900) and everything it calls (main.c:102-841).  The
 * reference has no plugin/FFI interface: the "API" of this path is two void functions with
 * external linkage, `sim_init(args_t)` (main.c:209) and `sim_step()` (main.c:843), plus the
 * ~20 file-scope globals they mutate (main.c:64-100, 552, 577-578) and that `draw_rows()`
 * reads back (main.c:921-943).  This header turns that seam into an explicit handle-based
 * C interface; INTEGRATION.md shows the patch a maintainer of the reference would apply.
 *
 * Conventions
 *   - plain C, no CUDA/torch types in any signature; pointers are HOST pointers unless a
 *     parameter says "device".  Host buffers are copied during the call, never retained.
 *   - every function returns 0 on success or a negative EULER_E_* code; the message of the
 *     last failure on the calling thread is `euler_gpu_last_error()`.  The library never
 *     calls exit() and lets no C++ exception cross the boundary (the reference exit(1)s /
 *     die()s instead: main.c:213-214, misc/terminal.c:48-52).
 *   - all planes are row-major [ny][nx], x fastest, exactly the reference's [Y][X] arrays
 *     (main.c:64).  U faces are valid for x < nx-1, V faces for y < ny-1 (main.c:36-43).
 *   - a handle is not thread-safe; one host thread drives it.  Work is enqueued on one CUDA
 *     stream; get/read/stats calls synchronise that stream.
 *   - there is NO CPU fallback: if no CUDA device is usable, create() fails.
 */
#ifndef EULER_GPU_H
#define EULER_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EULER_GPU_ABI_VERSION 4

enum euler_error {
  EULER_OK            =  0,
  EULER_E_INVALID     = -1,   /* bad argument */
  EULER_E_CUDA        = -2,   /* CUDA runtime error (message has the details) */
  EULER_E_NOMEM       = -3,
  EULER_E_UNSUPPORTED = -4,
  EULER_E_COMM        = -5    /* NCCL / multi-GPU plumbing */
};

/* Preconditioner of the pressure solve (replaces apply_preconditioner, main.c:580-627). */
enum euler_precon {
  EULER_PRECON_IC0_WAVEFRONT = 0, /* reference-faithful natural-order IC(0) ("MIC(0)-style",
                                     sigma=0.25, tau=0), solved by a pipelined anti-diagonal
                                     wavefront; reproduces the reference's iterates, incl.
                                     the stale-g_precon reads.  Single GPU. */
  EULER_PRECON_REDBLACK      = 1  /* GPU-parallel red-black ordered IC(0): same matrix, same
                                     tolerance, different (fully parallel) factor. */
};

/* How the PCG dot products (dot(), main.c:629-639) are summed. */
enum euler_dot_mode {
  EULER_DOT_TREE            = 0, /* parallel tree reduction (deterministic, fastest) */
  EULER_DOT_REFERENCE_ORDER = 1  /* the reference's sequential row-major sum, bit for bit.  With
                                    the IC(0) wavefront this makes the whole solve — alpha,
                                    beta, p, u, v — bit-identical to the reference's, even where
                                    the reference stops unconverged at its iteration cap and
                                    1-ulp differences would otherwise be amplified.  Latency-
                                    bound (one dependent add per cell). */
};

/* Storage precision of the PCG vectors (SURVEY §8b `pcg_dtype`, §8f row 4). */
enum euler_pcg_dtype {
  EULER_PCG_FP64 = 0, /* the reference's: every vector of project() is double (main.c:716-745). */
  EULER_PCG_FP32 = 1  /* NOT in the reference.  r, z, s, q, A s and the preconditioner diagonal
                         are stored and combined in fp32; the pressure p, the dot products and
                         alpha / beta / sigma stay fp64, and every `pcg_refresh_every` iterations
                         r is replaced by the true residual b - A p evaluated in fp64.  75 instead
                         of 132 B/cell per iteration.  Converged solves agree with the fp64 solve
                         to fp32 rounding of u, v (measured <= 1e-7 relative on the shipped
                         scenarios); a solve cut off at max_iterations is a different, equally
                         unconverged iterate.  Needs precon = REDBLACK, dot_mode = TREE,
                         stencil_variant = 0.  Works on slab handles (fp32 halo rows of z over
                         NVLink / NCCL).  CPU mirror: oracle/euler_oracle.c pcg_mixed. */
};

/* How the marker array is kept. */
enum euler_marker_mode {
  EULER_MARKERS_REFERENCE = 0, /* bit-identical ARRAY, not just multiset: reference swap-delete
                                  order (main.c:112) and the `dt -= t_prev` carry-over between
                                  successive markers (main.c:464,501,518) are reproduced. */
  EULER_MARKERS_FAST      = 1  /* every marker gets the full sub-step dt; order unspecified.
                                  Required for slab-decomposed (multi-GPU) runs. */
};

/* Planes / arrays addressable through euler_gpu_get/set (reference global in brackets). */
enum euler_field {
  EULER_F_U = 0,          /* float  [ny][nx]  g_u            main.c:64 */
  EULER_F_V,              /* float            g_v            main.c:65 */
  EULER_F_UTMP,           /* float            g_utmp         main.c:66 */
  EULER_F_VTMP,           /* float            g_vtmp         main.c:67 */
  EULER_F_SOLID,          /* uint8            g_solid        main.c:71 */
  EULER_F_SOURCE,         /* uint8            g_source       main.c:72 */
  EULER_F_SINK,           /* uint8            g_sink         main.c:73 */
  EULER_F_COUNT,          /* uint8            g_marker_count main.c:96  (== g_fluid) */
  EULER_F_PREV_COUNT,     /* uint8            g_prev_marker_count main.c:97 */
  EULER_F_MARKERS,        /* float  [n][2]    g_markers      main.c:95  (n = current length) */
  EULER_F_PRECON,         /* double           g_precon       main.c:577 (persistent) */
  EULER_F_Q,              /* double           g_q            main.c:578 */
  EULER_F_ADIAG,          /* int8             g_a            main.c:552 */
  EULER_F_P,              /* double           p (VLA in project, main.c:739) */
  EULER_F_R,              /* double           r / b          main.c:716,740 */
  EULER_F_Z,              /* double           z              main.c:743 */
  EULER_F_S,              /* double           s              main.c:745 */
  EULER_F_CR,             /* float            g_r            main.c:77  (only with params.rainbow) */
  EULER_F_CG,             /* float            g_g            main.c:78 */
  EULER_F_CB,             /* float            g_b            main.c:79 */
  /* fp32 twins of the PCG vectors (only with params.pcg_dtype = EULER_PCG_FP32) */
  EULER_F_R32,            /* float            r */
  EULER_F_Z32,            /* float            z = M^-1 r */
  EULER_F_S32,            /* float            s */
  EULER_F_Q32,            /* float            q (forward solve) / A s inside an iteration */
  EULER_F_PRECON32,       /* float            red-black preconditioner diagonal */
  EULER_F__COUNT
};

/* Stages runnable one at a time from the current device state (parity harness).
 * `dt` is the sub-step length where the stage takes one. */
enum euler_stage {
  EULER_S_ADVECT_MARKERS = 0, /* advect_markers(dt)            main.c:464-537 */
  EULER_S_REFRESH_COUNTS,     /* refresh_marker_counts()       main.c:102-117 */
  EULER_S_SOURCES,            /* update_fluid_sources()        main.c:276-298 */
  EULER_S_EXTRAPOLATE,        /* extrapolate(u),(v) + zero_bounds(u),(v)  main.c:865-868 */
  EULER_S_ADVECT_VELOCITY,    /* advect_u, advect_v, apply_body_forces, zero_bounds(utmp),(vtmp)
                                 main.c:871-889: (u,v) -> (utmp,vtmp) */
  EULER_S_PROJECT,            /* project(dt, utmp, vtmp, u, v) main.c:709-806 */
  EULER_S_BUILD_RHS,          /* b and a_diag only             main.c:713-733: -> R, ADIAG, P=0 */
  EULER_S_PRECONDITION,       /* z = M^-1 r                    main.c:580-627 */
  EULER_S_APPLY_A,            /* z = A s                       main.c:679-691 */
  EULER_S_PRESSURE_UPDATE,    /* clamp p, subtract gradient    main.c:769-805 */
  EULER_S_EXTRAPOLATE_COLOR,  /* --rainbow: extrapolate(g_r|g_g|g_b, P)   main.c:859-863 */
  EULER_S_ADVECT_COLOR,       /* --rainbow: advect_p x3 + plane copies    main.c:873-882 */
  EULER_S_FUSED_TAIL,         /* parity hook of the fused red-black iteration's second kernel (not a
                                 reference stage): with alpha = `dt`, from R = r, Q = A s, S = s, P = p
                                 and the current count plane: R <- r - alpha A s, P <- p + alpha s
                                 (main.c:753-754), Z <- M^-1 R (red-black IC(0)) in ONE launch */
  EULER_S__COUNT
};

typedef struct euler_params {
  /* physical constants; defaults = reference (main.c:58-60) */
  float h;               /* k_side_length = 1 */
  float rho;             /* k_density     = 1 */
  float gravity;         /* k_gravity     = -10 */
  /* sub-stepping; defaults = reference (main.c:838, 849-851) */
  float frame_time;      /* 0.1 s of simulated time per euler_gpu_step_frame */
  int   max_substeps;    /* 8 */
  float cfl_distance;    /* 0.75 (cells) */
  /* pressure solve; defaults = reference (main.c:735-736) */
  int    max_iterations; /* 100 */
  double tol;            /* (double)1e-6f, absolute, on ||r||inf */
  int    precon;         /* enum euler_precon; default IC0_WAVEFRONT */
  int    marker_mode;    /* enum euler_marker_mode; default REFERENCE */
  int    dot_mode;       /* enum euler_dot_mode; default TREE */
  /* source RNG: state of randf()'s xorshift64* stream (main.c:204) AFTER the host seeded
   * the initial markers; default = the reference seed 0x9bd185c449534b91 */
  uint64_t rng_state;
  /* plumbing */
  int   device;          /* CUDA device ordinal; default 0 */
  void *stream;          /* cudaStream_t to enqueue on; NULL = the library creates one */
  int   pcg_check_every; /* iterations enqueued between convergence polls; default 8 */
  int   stencil_variant; /* PCG kernels: 0 = TMA bulk-copy row pipeline, update_search fused
                            into apply_a (default); 1 = register sliding window, nothing
                            fused; 2 = as 0 plus axpy fused into the forward solve (kept for
                            A/B measurements) */
  /* row-slab decomposition: this handle owns global rows
   * [slab_row0, slab_row0 + slab_rows) of the nx x ny grid passed to create(); slab_rows = 0
   * means "not decomposed".  create() is given the GLOBAL planes and markers on every rank
   * and keeps its slab (+4 halo rows); get()/read_marker_count() write only the owned rows of
   * a global-shaped buffer.  Needs precon=REDBLACK, marker_mode=FAST.  See euler_gpu_comm_init. */
  int   slab_row0;
  int   slab_rows;
  /* --rainbow (main.c:76, 984-993): transport a passive RGB colour with the fluid — colorize()
   * at create/reinit (main.c:271-273), extrapolate(P) (:859-863), source colours (:292-294) and
   * advect_p (:873-882) every sub-step.  Six more fp32 planes.  Works on slab handles too (halo
   * rows of the three colour planes are exchanged once per sub-step).
   * Default 0, like the reference. */
  int   rainbow;
  /* mixed-precision PCG (enum euler_pcg_dtype); default EULER_PCG_FP64, like the reference */
  int   pcg_dtype;
  int   pcg_refresh_every; /* FP32 mode: iterations between residual replacements; even, or 0 =
                              never; default 10 */
} euler_params;

#define EULER_KERNEL_CLASSES 24

typedef struct euler_stats {
  uint64_t frames, substeps;        /* since create */
  uint64_t solves, solves_skipped;  /* project() calls that ran PCG / hit all_zero (main.c:742) */
  uint64_t pcg_iterations;          /* total */
  int      last_iterations;         /* of the last solve */
  double   last_residual;           /* ||r||inf at exit of the last solve */
  float    last_dt;
  uint64_t n_markers;
  int      source_exhausted;        /* g_source_exhausted, main.c:94 */
  uint64_t rng_state;
  uint64_t kernel_launches;         /* CUDA kernels launched by this handle since create */
  uint64_t device_bytes;            /* device memory owned by the handle */
  double   ms_markers, ms_grid, ms_project;  /* device time per stage group, summed, only
                                                when profiling was enabled */
  /* per kernel class (euler_gpu_kernel_class_name), only when profiling was enabled:
   * summed CUDA-event time and number of timed launches groups */
  uint64_t active_cells;            /* cells in PCG tiles that contain fluid (last solve): the
                                       cells the PCG kernels actually stream */
  double   kernel_ms[EULER_KERNEL_CLASSES];
  uint64_t kernel_count[EULER_KERNEL_CLASSES];
  uint64_t markers_migrated;        /* slab handles: markers this rank handed to a neighbouring slab
                                       since create (advect_markers moved them across the boundary) */
  uint64_t grid_cells;              /* cells the grid-stage kernels streamed in the last sub-step */
} euler_stats;

/* Size-independent invariants of the current state (euler_gpu_check): what a run on N row slabs
 * must reproduce from a run on one GPU.  Counted over the rows the handle OWNS, so the integer
 * fields of all ranks ADD UP (modulo 2^64) to the single-GPU values exactly, the sums add up to
 * within fp64 summation order (and whatever the solve itself differs by), the maxima combine by
 * max.  Reference quantities: g_markers_length (main.c:93), g_marker_count (main.c:96), the
 * divergence of main.c:720 evaluated on the projected u, v (SURVEY north_star: "matching
 * post-projection divergence norms"). */
typedef struct euler_check {
  uint64_t n_markers;       /* markers held by this handle */
  uint64_t fluid_cells;     /* owned cells with marker count > 0 */
  uint64_t count_sum;       /* sum of the owned count plane */
  uint64_t count_hash;      /* sum over owned cells of count * mix(global cell index), mod 2^64:
                               position-sensitive, additive over slabs */
  double   sum_abs_u;       /* sum |u| over owned rows */
  double   sum_abs_v;
  double   sum_p;           /* sum of p over owned fluid cells (p >= 0 after the clamp, main.c:773) */
  double   max_abs_div;     /* max over owned fluid cells of |(u - u[x-1] + v - v[y-1]) / h| */
  double   max_abs_u, max_abs_v;
} euler_check;

typedef struct euler_gpu euler_gpu;

/* Fill *p with the reference's constants. */
int euler_gpu_default_params(euler_params *p);

/* Replaces the state hand-over at the end of sim_init (main.c:209-274): the host keeps the
 * scenario parser and the marker seeding (so file format and RNG stream stay byte-identical)
 * and passes the three static masks, the seeded markers and the RNG state.  The ring of
 * sinks (main.c:244-252) must already be present in `sink`.  Runs refresh_marker_counts once,
 * like sim_init (main.c:268). */
int euler_gpu_create(euler_gpu **out, int nx, int ny,
                     const uint8_t *solid, const uint8_t *source, const uint8_t *sink,
                     const float *markers_xy, size_t n_markers,
                     const euler_params *params);
int euler_gpu_destroy(euler_gpu *h);
/* sim_init() again on an existing handle (same grid, same params): the hand-over of create()
 * without the allocation — new masks, markers and RNG state go to the device, every dynamic
 * plane (u, v, counts, the persistent g_precon, the PCG vectors) restarts from zero like the
 * reference's zero-initialised globals (main.c:64-100, 577), g_frame_count restarts at 0
 * (main.c:89), refresh_marker_counts runs once (main.c:268).  The launch / iteration statistics
 * keep counting.  On slab handles: collective (halo exchange of the count plane), after
 * euler_gpu_comm_init.
 * Slab handles (create and reinit): `markers_xy` may be ANY superset of the markers whose cell row
 * the slab owns — the whole global array, or just the slab's own markers (with markers stored in
 * row-major cell order that is one contiguous range of the global array, so each rank ships
 * 1/N of it over PCIe instead of all of it); the handle keeps the ones it owns.  The mask
 * pointers are global-shaped [ny][nx]; only the rows the handle stores are read. */
int euler_gpu_reinit(euler_gpu *h, const uint8_t *solid, const uint8_t *source, const uint8_t *sink,
                     const float *markers_xy, size_t n_markers, uint64_t rng_state);

/* sim_step() (main.c:843-900): up to max_substeps adaptive sub-steps covering frame_time.
 * Returns after the frame's work is complete on the device.  *substeps may be NULL. */
int euler_gpu_step_frame(euler_gpu *h, int *substeps);
/* calculate_timestep(frame_time) (main.c:834-841) from the current u, v. */
int euler_gpu_calculate_timestep(euler_gpu *h, float frame_time, float *dt);
/* One sub-step with the given dt: body of the loop main.c:855-893. */
int euler_gpu_substep(euler_gpu *h, float dt);
/* One stage (parity harness). */
int euler_gpu_run_stage(euler_gpu *h, int stage, float dt);

/* What draw_rows() reads every frame (main.c:933): the uint8 marker-count plane. */
int euler_gpu_read_marker_count(euler_gpu *h, uint8_t *dst /* ny*nx */);

/* The renderer's feed for large grids: draw_rows() only looks at the part of the grid that fits
 * the terminal window — rows [max(ny-1-g_wy, 1), ny-1), columns [1, min(nx-1, g_wx+1))
 * (main.c:917-920).  Copies the rectangle [x0, x0+w) x [y0, y0+hh) of a plane (any EULER_F_*
 * plane field) into the same position of a GLOBAL-shaped [ny][nx] host buffer and leaves the
 * rest of the buffer untouched: a 16384^2 run ships kilobytes per drawn frame instead of the
 * 268 MB plane.  Slab handles write the part of the window they own. */
int euler_gpu_read_window(euler_gpu *h, int field, int x0, int y0, int w, int hh, void *dst_global);

/* State access (checkpoint / parity).  `bytes` must equal the field's size: ny*nx*sizeof(T),
 * or n_markers*8 for EULER_F_MARKERS (set: any n <= 4*nx*ny; sets the length too). */
int euler_gpu_get(euler_gpu *h, int field, void *dst, size_t bytes);
int euler_gpu_set(euler_gpu *h, int field, const void *src, size_t bytes);
/* colorize() (main.c:187-201): what the `r` key does with --rainbow (main.c:971-974).  Error on
 * a handle without params.rainbow. */
int euler_gpu_colorize(euler_gpu *h);
int euler_gpu_set_rng_state(euler_gpu *h, uint64_t state);
int euler_gpu_set_source_exhausted(euler_gpu *h, int exhausted);
/* g_frame_count (main.c:89, incremented per euler_gpu_step_frame; read only by the source
 * colours of --rainbow, main.c:283): restored by a checkpoint load. */
int euler_gpu_set_frame_count(euler_gpu *h, uint64_t frames);
/* `max_iterations` of project() (main.c:735) for the solves that follow; 0 = rhs, p = 0 and the
 * velocity update only (every stage except the PCG iteration: what is bit-determined whatever the
 * order of the dot products — used by the cross-N invariants of bench.py and the parity tests). */
int euler_gpu_set_max_iterations(euler_gpu *h, int max_iterations);

int euler_gpu_stats(euler_gpu *h, euler_stats *out);
/* Invariants of the current state, reduced on the device (one pass over the owned rows; a few
 * hundred bytes come back).  Synchronises the stream. */
int euler_gpu_check(euler_gpu *h, euler_check *out);
/* Toggle per-stage-group CUDA-event timing (adds synchronisation; off by default). */
int euler_gpu_set_profiling(euler_gpu *h, int enabled);
/* Name of kernel class i (0 <= i < EULER_KERNEL_CLASSES), or NULL. */
const char *euler_gpu_kernel_class_name(int i);
/* Zero the profiling accumulators. */
int euler_gpu_reset_profile(euler_gpu *h);
/* Diagnostics (the new build's counterpart of the reference's debug helpers, misc/debug.c): with
 * EULER_TRACE=<slots> in the environment when the handle is created, the two kernels of a red-black
 * PCG iteration record a timeline per launch — 16 words: %globaltimer ns as (~min, max) over the
 * blocks for {block start, scalars known, rows done, block exit}, word 8 = kernel (1 search+apply,
 * 2 tail), word 9 = iteration (slab runs).  Copies the recorded slots out (up to max_slots) and
 * starts over; *n_slots = 0 when tracing is off.  With EULER_TRACE_BLOCKS=<kernel> and room for
 * 1024 more slots in `out`, the per-block records {start ns, rows-done ns, SM id, 0} x 4096 of that
 * kernel's last launch follow the slots. */
int euler_gpu_trace_read(euler_gpu *h, unsigned long long *out, size_t max_slots, size_t *n_slots);
int euler_gpu_synchronize(euler_gpu *h);
/* The CUDA stream (cudaStream_t) the handle enqueues on, for event timing by the caller. */
void *euler_gpu_stream(euler_gpu *h);

/* Benchmark hook: run exactly `iterations` PCG iterations of the current system without
 * convergence polling (used to time the iteration kernels alone; state must hold a built
 * rhs). */
int euler_gpu_pcg_iterations(euler_gpu *h, int iterations);

/* Multi-GPU (row slabs, one process per GPU).  `unique_id` is the 128-byte ncclUniqueId
 * produced on rank 0 by euler_gpu_comm_unique_id and distributed by the caller (any side
 * channel: torch.distributed, MPI, a file).  Collective over all ranks. */
int euler_gpu_comm_unique_id(void *unique_id_128);
/* Optional NVLink fast path for the per-iteration exchanges of the slab solve (two scalar
 * all-gathers and one halo exchange per PCG iteration): CUDA-IPC mapped peer memory and
 * hand-written store/flag kernels instead of NCCL (~5 us instead of ~45 us per exchange).
 * After comm_init every rank calls export, the caller all-gathers the 256-byte blobs in rank
 * order (any side channel) and every rank calls import.  Collective.  Without it the NCCL
 * path is used. */
#define EULER_P2P_BLOB_BYTES 256
int euler_gpu_comm_p2p_export(euler_gpu *h, void *blob_256);
int euler_gpu_comm_p2p_import(euler_gpu *h, const void *blobs /* n_ranks * 256 bytes */);
/* Balanced contiguous split of global_ny rows over n_ranks (pure host arithmetic). */
int euler_gpu_slab_partition(int global_ny, int n_ranks, int rank, int *row0, int *rows);
/* Same, but balancing the sum of row_weight[y] (e.g. fluid cells per row + a small constant:
 * the PCG only streams tiles that contain fluid) instead of the row count. */
int euler_gpu_slab_partition_weighted(const uint64_t *row_weight, int global_ny, int n_ranks,
                                      int rank, int *row0, int *rows);
int euler_gpu_comm_init(euler_gpu *h, int rank, int n_ranks, const void *unique_id_128);

const char *euler_gpu_last_error(void);
int euler_gpu_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* EULER_GPU_H */
