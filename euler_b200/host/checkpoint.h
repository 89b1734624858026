/* euler_b200/host/checkpoint.h — save / restore of the simulation state through the C-ABI's
 * get/set surface.  The reference keeps its state in file-scope globals and has
 * no checkpoint; what must survive is exactly what sim_step() reads at its start:
 * g_u, g_v (main.c:64-65), g_marker_count / g_prev_marker_count (:96-97), g_markers +
 * g_markers_length (:93-95), g_source_exhausted (:94), the RNG state of randf() (:204), the
 * persistent g_precon plane (:577), g_frame_count (:89) and, with --rainbow,
 * g_r g_g g_b (:77-79).  Everything else is recomputed every sub-step.
 *
 * File: "EULERCK1" | nx ny (i32) | flags (u32, bit 0 = colour planes, bit 1 = the handle had no
 * fp64 g_precon plane (pcg_dtype = FP32): zeros in its place) | n_markers (u64) |
 * rng_state (u64) | frames (u64) | source_exhausted (i32) | pad (i32) | u v (f32 planes) |
 * count prev_count (u8 planes) | precon (f64 plane) | [r g b (f32 planes)] | markers (f32 x 2).
 * Little-endian, planes row-major [ny][nx]. */
#ifndef EULER_CHECKPOINT_H
#define EULER_CHECKPOINT_H
#include "euler_gpu.h"

typedef struct euler_ckpt_api {   /* the entry points of libeuler_gpu.so the two calls need */
  int (*get)(euler_gpu *, int, void *, size_t);
  int (*set)(euler_gpu *, int, const void *, size_t);
  int (*stats)(euler_gpu *, euler_stats *);
  int (*set_rng_state)(euler_gpu *, uint64_t);
  int (*set_source_exhausted)(euler_gpu *, int);
  int (*set_frame_count)(euler_gpu *, uint64_t);
} euler_ckpt_api;

/* 0 ok; -1 I/O or allocation error; -2 the file does not fit the handle (size, colour planes);
 * > 0: an EULER_E_* code from the library, negated. */
int euler_checkpoint_save(const euler_ckpt_api *a, euler_gpu *sim, int nx, int ny, int rainbow, const char *path);
int euler_checkpoint_load(const euler_ckpt_api *a, euler_gpu *sim, int nx, int ny, int rainbow, const char *path);
#endif
