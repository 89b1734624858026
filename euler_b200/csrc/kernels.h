// euler_b200/csrc/kernels.h — host-side launchers of the sm_100a kernels (one per stage of
// reference sim_step(), main.c:843-900).  Every launcher enqueues on ctx.stream and bumps
// ctx.launches by the number of kernels it started.
#pragma once
#include "common.cuh"
#include "interp.cuh"
#include "p2p.cuh"

namespace euler {

// Kernel classes for the per-kernel timers (euler_stats.kernel_ms); order is ABI.
enum KernelClass {
  KC_MAXSQ = 0, KC_ADVECT_MARKERS, KC_REFRESH_COUNTS, KC_SOURCES, KC_EXTRAPOLATE,
  KC_ADVECT_VELOCITY, KC_BUILD_RHS, KC_PRECON_BUILD, KC_PRECON_APPLY, KC_APPLY_A, KC_AXPY,
  KC_UPDATE_SEARCH, KC_PRESSURE_UPDATE, KC_MISC, KC_FUSED_A, KC_FUSED_B,
  KC_PRECON_FWD, KC_PRECON_BWD, KC_COLOR, KC_RESIDUAL, KC_FUSED_TAIL, KC__COUNT
};

struct Prof {
  bool on;
  int n, cap;
  cudaEvent_t* ev;     // 2*cap events
  int* cls;            // class of each recorded pair
  double ms[KC__COUNT];
  unsigned long long count[KC__COUNT];
};

struct Ctx {
  Prof prof;
  Grid g;                         // all locally stored rows
  int own0, own1;                 // rows of g this handle owns (== 0, g.ny when not decomposed)
  InterpLimits lim;
  cudaStream_t stream;
  unsigned long long launches;
  int sm_count;
  // constants (reference main.c:58-60, 735-736, 838)
  float h, rho, gravity;
  int dot_mode;                   // 0 tree reduction, 1 reference-order sequential sum
  int use_pipe;                   // stencil kernels: 1 TMA row pipeline, 0 register window
  int distributed;                // row-slab mode: reductions are finished across ranks (comm.cu)
  // static masks
  uint8_t *solid, *source, *sink;
  // dynamic cell classification: marker counts now / previous sub-step (main.c:96-97)
  uint8_t *count, *prev_count;
  unsigned int* count32;          // atomic binning target, folded to uint8 afterwards
  // velocities (main.c:64-67) + one scratch pair for the out-of-place extrapolation
  float *u, *v, *utmp, *vtmp, *uext, *vext;
  // --rainbow colour planes (main.c:77-82); null unless euler_params.rainbow
  float *cr, *cg, *cb, *crtmp, *cgtmp, *cbtmp;
  // markers (main.c:92-95): ping-pong AoS float2 arrays
  float2 *markers, *markers_alt;
  size_t max_markers;             // capacity of the local arrays
  size_t max_markers_global;      // MAX_MARKER_COUNT of the whole grid (main.c:92)
  // compaction scratch
  unsigned int* seg_count;        // per 1024-marker segment
  unsigned int* seg_offset;
  void* cand;                     // reference marker mode: ordered rewind records
  float* cand_dt;                 //   and dt after each of them
  size_t cand_cap;
  size_t n_segments;
  // source cells, row-major (static)
  unsigned int* source_cells;
  size_t n_source_cells;
  size_t n_source_cells_global;
  unsigned long long* rng_jump;   // 64 x 64 columns of T^(2^j), xorshift64 transition
  // pressure solve
  int8_t* adiag;
  double *precon, *q, *p, *r, *z, *s;
  double *s2, *r2;                // twins of s and r for the fused red-black iteration
  // mixed-precision mode (euler_params.pcg_dtype = FP32): fp32 twins of r, z, s (two planes,
  // ping-pong), q and the preconditioner diagonal; p stays fp64 and the fp64 r plane keeps b
  int mixed;
  float *r32, *z32, *s32, *s32b, *q32, *pc32;
  int ns_mixed[3];                // TMA ring depth of the fp32 forward / backward / search+apply kernels
  int mixed_blocks;               // 8: the 64-register fp32 instantiations (8 resident blocks per SM); 0: default
  int fused;                      // red-black: two fused kernels per iteration
  uint8_t* tile_active;           // per PCG tile: contains fluid (pcg_kernels.cu)
  int* tile_list;                 // ordered compact list of those tiles
  double* partials;               // grid-reduction scratch
  size_t n_partials;
  unsigned int* wf_progress;      // wavefront strip progress flags
  int n_strips;
  // tiles the grid stages stream (common.cuh GridTiles): "has fluid" flags of this and the two
  // previous sub-steps (ring), the compact list, how many previous flag planes are valid, and
  // whether this sub-step's launches use the list
  uint8_t* gt_flags[3];
  int* gt_list;
  int gt_tx, gt_ty;
  int gt_hist;                    // valid previous flag planes (0..2); the list needs 2
  int gt_sparse;                  // this sub-step's grid stages go by the list
  int gt_prev_sparse;             // last sub-step's list is valid (the count fold uses it)
  DevScalars* sc;                 // device scalars
  // row slabs, NVLink path (p2p.cuh): 0 = exchanges by NCCL, 1 = separate exchange kernels,
  // 2 = exchanges fused into the producing kernels' epilogues (default when peers are mapped)
  int p2p_mode;
  DistArgs dist;                  // z_dn / z_up here are row 0 of the neighbours' planes
  int p2p_dn_own1, p2p_up_own0;   // the neighbours' owned-row bounds in their local rows
  double tol;                     // PCG stop tolerance (main.c:736)
  // in-kernel timeline of the iteration kernels (common.cuh trace_mark): EULER_TRACE=<slots>
  unsigned long long* trace;      // trace_cap slots of TRACE_WORDS words, or null
  int trace_cap, trace_n;
};
// next free trace slot (null when tracing is off or the buffer is full)
inline unsigned long long* trace_slot(Ctx& c) {
  if (!c.trace || c.trace_n >= c.trace_cap) return nullptr;
  return c.trace + (size_t)TRACE_WORDS * (size_t)c.trace_n++;
}

// RAII: when profiling is on, brackets the launches of one kernel class with CUDA events on
// ctx.stream; prof_collect() (after a stream sync) folds them into Prof::ms.
struct ProfScope {
  Ctx& c; int slot;
  ProfScope(Ctx& ctx, int cls) : c(ctx), slot(-1) {
    Prof& p = c.prof;
    if (p.on && p.n < p.cap) { slot = p.n++; p.cls[slot] = cls; cudaEventRecord(p.ev[2 * slot], c.stream); }
  }
  ~ProfScope() { if (slot >= 0) cudaEventRecord(c.prof.ev[2 * slot + 1], c.stream); }
};
inline void prof_collect(Ctx& c) {
  Prof& p = c.prof;
  for (int i = 0; i < p.n; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.ev[2 * i], p.ev[2 * i + 1]) == cudaSuccess) {
      p.ms[p.cls[i]] += ms; p.count[p.cls[i]] += 1;
    }
  }
  p.n = 0;
}

// ---- grid stages (grid_kernels.cu)
void launch_grid_tiles(Ctx& c);                              // count plane -> the tile list of this sub-step
inline GridTiles grid_tiles_of(const Ctx& c, bool sparse) {
  GridTiles gt;
  gt.list = sparse ? c.gt_list : nullptr;
  gt.count = &c.sc->grid_tiles;
  gt.tx = c.gt_tx; gt.ty = c.gt_ty;
  return gt;
}
void launch_maxsq(Ctx& c);                                   // -> sc.max_u2_bits/max_v2_bits
void launch_timestep(Ctx& c, float frame_time, float cfl);   // -> sc.dt (uses sc.max_*)
void launch_extrapolate(Ctx& c);                             // (u,v) -> (uext,vext); caller swaps
void launch_advect_velocity(Ctx& c, float dt);               // (u,v) -> (utmp,vtmp)
void launch_build_rhs(Ctx& c, float dt);                     // -> r, p=0, adiag, sc.nonzero_rhs
void launch_pressure_update(Ctx& c, float dt);               // p,utmp,vtmp -> u,v (+max u2,v2)
// euler_gpu_check: invariants over the owned rows; ints3 = {fluid cells, sum count, count hash},
// parts = 6 doubles per block {sum|u|, sum|v|, sum p, max|div|, max|u|, max|v|}
void launch_check(Ctx& c, unsigned long long* ints3, double* parts, int blocks);
// --rainbow colour transport; all four return at once when the colour planes do not exist
void launch_colorize(Ctx& c);                                // colorize(), main.c:187-201
void launch_extrapolate_color(Ctx& c);                       // extrapolate(r|g|b, P), main.c:859-863
void launch_source_colors(Ctx& c, unsigned int frame_count); // colour writes of main.c:292-294
void launch_advect_color(Ctx& c, float dt);                  // advect_p x3 + copies, main.c:873-882

// ---- markers (marker_kernels.cu)
void launch_advect_markers(Ctx& c, float dt, int mode);      // in: markers, out: markers (swapped inside)
void launch_refresh_counts(Ctx& c);                          // prev<-cur, re-bin, delete in sink/solid
void launch_sources(Ctx& c);                                 // update_fluid_sources
void launch_sources_count(Ctx& c);
// sim_init: the ordered list of source cells from the uploaded source plane.  rows_scratch: g.ny
// words; totals2[0] = cells in owned rows (the list), [1] = in all stored rows
void launch_source_rows_count(Ctx& c, unsigned int* rows_scratch, unsigned long long* totals2);
void launch_source_rows_write(Ctx& c, const unsigned int* rows_scratch);
void launch_sources_prep(Ctx& c, const double* gathered, int rank, int nranks);
// slab mode: advect + hand the markers that left rows [own_lo, own_hi) to the staging buffers
// (null = no neighbour on that side); they are deleted locally by the next refresh_marker_counts
void launch_advect_markers_slab(Ctx& c, float dt, int own_lo_global, int own_hi_global, float2* send_dn,
                                float2* send_up, size_t send_cap);
void launch_add_markers(Ctx& c, unsigned long long n);       // sc.n_markers += n (markers received)
void launch_partition_markers(Ctx& c, int own_lo_global, int own_hi_global, float2* send_dn,
                              float2* send_up, size_t send_cap, unsigned long long* n_keep);
// slab create/reinit: append the staged markers of rows [lo, hi) to c.markers at sc->n_markers
void launch_filter_markers(Ctx& c, const float2* staged, size_t n, int own_lo_global, int own_hi_global);
void init_rng_jump_table(unsigned long long* host_table /* 64*64 */);
size_t marker_candidate_bytes();

// ---- pressure solve (pcg_kernels.cu, wavefront.cu)
void launch_ic0_build(Ctx& c);                               // E^-1, wavefront (once per project)
void launch_ic0_apply(Ctx& c, bool init);                    // z = M^-1 r (+ z.r, sigma/beta)
void launch_rb_build(Ctx& c);                                // red-black E^-1
void launch_rb_apply(Ctx& c, bool init);                     // z = M^-1 r (+ z.r, sigma/beta)
void launch_rb_forward(Ctx& c);                              //   q = L^-1 r
void launch_rb_backward(Ctx& c, bool init);                  //   z = L^-T q (+ z.r)
// split_it: 0, or the 1-based iteration number when the slab solve uses the split-phase exchange
void launch_fused_search_apply(Ctx& c, bool init, int split_it = 0);   // s' = z + beta s ; A s' ; alpha
void launch_fused_axpy_forward(Ctx& c, double tol);           // p, r', ||r'||inf, q = L^-1 r'
// r' = r - alpha A s, p, ||r'||inf, q = L^-1 r', z = L^-T q, z.r', beta: one kernel (pcg_tail.cuh);
// mode as launch_axpy
void launch_fused_tail(Ctx& c, double tol, int mode, int split_it = 0);
void launch_dist_peek(Ctx& c, bool apply);                   // split-phase: the pending {z.r, ||r||inf} for the host
void launch_set_alpha(Ctx& c, double alpha);                 // parity hook: alpha = given, alpha_prev = 0, sigma = 1
void launch_dist_alpha(Ctx& c, const double* gathered, int nranks);
void launch_dist_beta(Ctx& c, const double* gathered, int nranks, bool init, double tol);
void launch_copy_search(Ctx& c);                             // s = z
void launch_apply_a(Ctx& c, bool with_alpha);                // z = A s (+ z.s, alpha)
// mode 2: p += a s, r -= a (A s), ||r||inf every iteration (the reference's order of updates);
// fused iteration: mode 0 on odd iterations (r only), mode 1 on even ones (r and both pending
// p updates) — see k_axpy
void launch_axpy(Ctx& c, double tol, bool as_in_q = false, int mode = 2);
// completes p after a fused solve that stopped on an odd iteration; s_odd_plane = the plane
// the first iteration wrote its search direction to (Ctx::s2 at the start of the solve)
void launch_p_fixup(Ctx& c, const void* s_odd_plane, int split = 0);
// mixed-precision mode: r32 <- b - A p evaluated in fp64 (residual replacement)
void launch_true_residual(Ctx& c);
void launch_update_search(Ctx& c);                           // s = z + beta s
void launch_pcg_reset(Ctx& c);                               // iters=0, done=0
void launch_tile_flags(Ctx& c);                              // per-tile fluid flags from count
int pcg_tile_count(const Grid& g);
void launch_dot_zr_exact(Ctx& c, bool init);                 // no-op unless dot_mode
int pcg_tile_cells(const Ctx& c);

}  // namespace euler
