// euler_b200/csrc/api.cu — the C-ABI of libeuler_gpu.so (include/euler_gpu.h): handle,
// device memory, state access and the host-side drivers of sim_step() / project()
// (reference main.c:843-900, 709-806).  No torch, no CPU fallback: every stage is a CUDA
// kernel from grid_kernels.cu / marker_kernels.cu / pcg_kernels.cu / wavefront.cu.
#include "../../include/euler_gpu.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include <vector>

#include "comm.h"
#include "kernels.h"

using namespace euler;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return fail(EULER_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                 \
  } while (0)

}  // namespace

constexpr int SLAB_HALO = 4;     // halo rows kept on each side of the owned rows
constexpr int NS_MIXED_DEFAULT[3] = {4, 4, 4};   // forward, backward, search+apply (pcg_dtype = FP32)
static_assert(SLAB_HALO == P2P_HALO_DEPTH, "p2p.cuh halo depth");

struct euler_gpu {
  Ctx c;
  euler_params prm;
  int nx, ny;                    // GLOBAL grid size
  bool slab;                     // row-slab decomposition configured (params.slab_rows > 0)
  int row0, rows, lo;            // first owned global row, owned rows, first stored global row
  Comm cm;
  bool comm_ready;
  unsigned long long* n_keep;    // device scratch of the marker partition
  size_t source_cap;             // capacity of Ctx::source_cells
  unsigned int* src_rows;        // per stored row: source cells, then their offset in the list (load_state)
  unsigned long long* src_totals;
  cudaStream_t up_stream;        // sim_init hand-over: uploads next to the zeroing of the planes
  cudaEvent_t up_event;
  P2P pp;                        // NVLink peer-to-peer fast path of the per-iteration exchanges
  void* z_raw;                   // cudaMalloc base of the z plane (for the IPC handle)
  bool own_stream;
  std::vector<void*> allocs;     // raw cudaMalloc pointers
  DevScalars* host_sc;           // pinned mirror
  size_t device_bytes;
  bool max_valid;                // sc.max_*_bits describe the current u, v
  bool use_tail;                 // fused red-black iteration: axpy + forward + backward as one kernel
  bool split;                    // slab solve over NVLink: split-phase scalar exchange (p2p.cuh)
  int poll_hint;                 // iterations of the previous solve (first_poll_chunk)
  // stats
  uint64_t frames, substeps, solves, solves_skipped, pcg_iterations, markers_migrated;
  int last_iterations;
  double last_residual;
  float last_dt;
  bool profiling;
  cudaEvent_t ev[4];
  double ms_markers, ms_grid, ms_project;
};

namespace {

template <class T>
int alloc_plane(euler_gpu* h, T** out) {
  const Grid& g = h->c.g;
  const size_t rows = (size_t)g.ny + 2 * GUARD_ROWS;
  const size_t bytes = rows * g.pitch * sizeof(T);
  void* raw = nullptr;
  CU(cudaMalloc(&raw, bytes));
  CU(cudaMemsetAsync(raw, 0, bytes, h->c.stream));
  h->allocs.push_back(raw);
  h->device_bytes += bytes;
  *out = reinterpret_cast<T*>(raw) + (size_t)GUARD_ROWS * g.pitch;
  return 0;
}

template <class T>
int alloc_array(euler_gpu* h, T** out, size_t n) {
  void* raw = nullptr;
  const size_t bytes = (n ? n : 1) * sizeof(T);
  CU(cudaMalloc(&raw, bytes));
  CU(cudaMemsetAsync(raw, 0, bytes, h->c.stream));
  h->allocs.push_back(raw);
  h->device_bytes += bytes;
  *out = reinterpret_cast<T*>(raw);
  return 0;
}

int pull_scalars(euler_gpu* h) {
  CU(cudaMemcpyAsync(h->host_sc, h->c.sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, h->c.stream));
  CU(cudaStreamSynchronize(h->c.stream));
  return 0;
}

struct FieldInfo { void* ptr; size_t elem; };

FieldInfo field_info(euler_gpu* h, int f) {
  Ctx& c = h->c;
  switch (f) {
    case EULER_F_U: return {c.u, 4};
    case EULER_F_V: return {c.v, 4};
    case EULER_F_UTMP: return {c.utmp, 4};
    case EULER_F_VTMP: return {c.vtmp, 4};
    case EULER_F_SOLID: return {c.solid, 1};
    case EULER_F_SOURCE: return {c.source, 1};
    case EULER_F_SINK: return {c.sink, 1};
    case EULER_F_COUNT: return {c.count, 1};
    case EULER_F_PREV_COUNT: return {c.prev_count, 1};
    case EULER_F_PRECON: return {c.precon, 8};
    case EULER_F_Q: return {c.q, 8};
    case EULER_F_ADIAG: return {c.adiag, 1};
    case EULER_F_P: return {c.p, 8};
    case EULER_F_R: return {c.r, 8};
    case EULER_F_Z: return {c.z, 8};
    case EULER_F_S: return {c.s, 8};
    case EULER_F_CR: return {c.cr, 4};
    case EULER_F_CG: return {c.cg, 4};
    case EULER_F_CB: return {c.cb, 4};
    case EULER_F_R32: return {c.r32, 4};
    case EULER_F_Z32: return {c.z32, 4};
    case EULER_F_S32: return {c.s32, 4};
    case EULER_F_Q32: return {c.q32, 4};
    case EULER_F_PRECON32: return {c.pc32, 4};
    default: return {nullptr, 0};
  }
}

// host arrays always have the GLOBAL shape [ny][nx]; the handle takes the rows it stores
int upload_plane(euler_gpu* h, void* dst, const void* src, size_t elem, cudaStream_t stream) {
  const Grid& g = h->c.g;
  const char* from = reinterpret_cast<const char*>(src) + (size_t)h->lo * g.nx * elem;
  CU(cudaMemcpy2DAsync(dst, g.pitch * elem, from, (size_t)g.nx * elem, (size_t)g.nx * elem, g.ny,
                       cudaMemcpyHostToDevice, stream));
  return 0;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(EULER_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

void prof_mark(euler_gpu* h, int i) {
  if (h->profiling) cudaEventRecord(h->ev[i], h->c.stream);
}


template <class T>
int zero_plane(euler_gpu* h, T* plane) {
  if (!plane) return 0;
  const Grid& g = h->c.g;
  const size_t rows = (size_t)g.ny + 2 * GUARD_ROWS;
  CU(cudaMemsetAsync(plane - (size_t)GUARD_ROWS * g.pitch, 0, rows * g.pitch * sizeof(T), h->c.stream));
  return 0;
}

// Slab handles scan only their own rows for source cells; the sources step is a collective
// (cross-rank prefix of the RNG draws), so every rank must know whether ANY rank has one.
int agree_on_sources(euler_gpu* h) {
  Ctx& c = h->c;
  int flag = c.n_source_cells_global ? 1 : 0;
  CU(cudaMemcpyAsync(&c.sc->pad1, &flag, sizeof flag, cudaMemcpyHostToDevice, c.stream));
  if (comm_allreduce_max_i32(c, h->cm, &c.sc->pad1, 1)) return fail(EULER_E_COMM, "%s", comm_last_error());
  CU(cudaMemcpyAsync(&flag, &c.sc->pad1, sizeof flag, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  if (flag && !c.n_source_cells_global) c.n_source_cells_global = 1;   // used as a boolean on slabs
  return 0;
}

// --rainbow on slabs: halo rows of the three colour planes from the neighbours' owned rows
int color_halos(euler_gpu* h) {
  Ctx& c = h->c;
  if (!c.cr) return 0;
  if (comm_halo(c, h->cm, c.cr, 4, SLAB_HALO) || comm_halo(c, h->cm, c.cg, 4, SLAB_HALO) ||
      comm_halo(c, h->cm, c.cb, 4, SLAB_HALO))
    return fail(EULER_E_COMM, "%s", comm_last_error());
  return 0;
}

// The state hand-over at the end of sim_init (main.c:209-274) into an existing handle: the
// three static masks, the seeded markers and the RNG state go to the device, every dynamic
// plane starts from zero (the reference's globals are zero-initialised: main.c:64-100, 577),
// and refresh_marker_counts runs once (main.c:268).  `fresh`: the planes were just allocated
// (already zero).  Host buffers may be pageable or pinned; nothing is retained.
int load_state(euler_gpu* h, const uint8_t* solid, const uint8_t* source, const uint8_t* sink,
               const float* markers_xy, size_t n_markers, uint64_t rng_state, bool fresh) {
  Ctx& c = h->c;
  // The uploads go out on a second stream, next to the zeroing of the dynamic planes on the main one
  // (copy engine and SMs: at 16384^2 the memsets are 39 GB, the uploads 2.5 GB over PCIe).  They
  // start after everything already enqueued on the main stream (a fresh handle's allocation memsets
  // clear the very planes the uploads fill).
  if (!h->up_stream) {
    CU(cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->up_event, cudaEventDisableTiming));
  }
  CU(cudaEventRecord(h->up_event, c.stream));
  CU(cudaStreamWaitEvent(h->up_stream, h->up_event, 0));
  if (!fresh) {
    int rc = 0;
    rc |= zero_plane(h, c.u); rc |= zero_plane(h, c.v); rc |= zero_plane(h, c.utmp); rc |= zero_plane(h, c.vtmp);
    rc |= zero_plane(h, c.uext); rc |= zero_plane(h, c.vext);
    rc |= zero_plane(h, c.count); rc |= zero_plane(h, c.prev_count); rc |= zero_plane(h, c.count32);
    rc |= zero_plane(h, c.adiag); rc |= zero_plane(h, c.precon); rc |= zero_plane(h, c.q);
    rc |= zero_plane(h, c.p); rc |= zero_plane(h, c.r); rc |= zero_plane(h, c.z); rc |= zero_plane(h, c.s);
    rc |= zero_plane(h, c.s2); rc |= zero_plane(h, c.r2);
    rc |= zero_plane(h, c.r32); rc |= zero_plane(h, c.z32); rc |= zero_plane(h, c.s32);
    rc |= zero_plane(h, c.s32b); rc |= zero_plane(h, c.q32); rc |= zero_plane(h, c.pc32);
    rc |= zero_plane(h, c.cr); rc |= zero_plane(h, c.cg); rc |= zero_plane(h, c.cb);
    rc |= zero_plane(h, c.crtmp); rc |= zero_plane(h, c.cgtmp); rc |= zero_plane(h, c.cbtmp);
    if (rc) return rc;
    h->max_valid = false;
  }
  int rc;
  if ((rc = upload_plane(h, c.solid, solid, 1, h->up_stream))) return rc;
  if ((rc = upload_plane(h, c.source, source, 1, h->up_stream))) return rc;
  if ((rc = upload_plane(h, c.sink, sink, 1, h->up_stream))) return rc;

  // Every dynamic plane is zero now, so "zero outside the tiles that hold or border fluid" — what the
  // tile list of the grid stages relies on (common.cuh GridTiles) — holds from the first sub-step:
  // the two "previous sub-step" flag planes start empty instead of forcing two passes over all tiles.
  // (The first count fold still runs over all tiles: gt_prev_sparse.)
  for (int i = 0; i < 3; ++i)
    CU(cudaMemsetAsync(c.gt_flags[i], 0, (size_t)c.gt_tx * c.gt_ty, c.stream));
  c.gt_hist = 2;
  c.gt_sparse = c.gt_prev_sparse = 0;
  memset(h->host_sc, 0, sizeof(DevScalars));
  h->host_sc->n_markers = h->slab ? 0 : n_markers;
  h->host_sc->rng_state = rng_state;
  if (c.trace) {
    const char* tk = getenv("EULER_TRACE_BLOCKS");      // 1 search+apply, 2 tail
    h->host_sc->trace_blk = c.trace + (size_t)c.trace_cap * TRACE_WORDS;
    h->host_sc->trace_kind = tk ? atoi(tk) : 0;
  }
  CU(cudaMemcpyAsync(c.sc, h->host_sc, sizeof(DevScalars), cudaMemcpyHostToDevice, c.stream));
  if (!h->slab) {
    if (n_markers > c.max_markers) return fail(EULER_E_INVALID, "too many markers");
    if (n_markers)
      CU(cudaMemcpyAsync(c.markers, markers_xy, n_markers * sizeof(float2), cudaMemcpyHostToDevice, h->up_stream));
  }
  CU(cudaEventRecord(h->up_event, h->up_stream));
  CU(cudaStreamWaitEvent(c.stream, h->up_event, 0));
  if (h->slab) {
    // keep the markers whose cell row this slab owns: the global array streams through the
    // scratch marker array in chunks and a kernel appends the owned ones (order is free in
    // FAST marker mode, the only one slabs support)
    const size_t chunk = c.max_markers;
    for (size_t off = 0; off < n_markers; off += chunk) {
      const size_t n = n_markers - off < chunk ? n_markers - off : chunk;
      CU(cudaMemcpyAsync(c.markers_alt, markers_xy + 2 * off, n * sizeof(float2), cudaMemcpyHostToDevice, c.stream));
      launch_filter_markers(c, c.markers_alt, n, h->row0, h->row0 + h->rows);
    }
  }

  // static row-major list of source cells (main.c:284-286 visits them in this order), built on the
  // device from the uploaded plane (marker_kernels.cu k_src_row_*).  A slab handle lists the rows it
  // owns and counts the rows it stores; whether ANY rank has a source is agreed on below.
  launch_source_rows_count(c, h->src_rows, h->src_totals);
  unsigned long long tot[2] = {0, 0}, kept = 0;
  CU(cudaMemcpyAsync(tot, h->src_totals, sizeof tot, cudaMemcpyDeviceToHost, c.stream));
  if (h->slab) CU(cudaMemcpyAsync(&kept, &c.sc->n_markers, sizeof kept, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));      // also: pageable host buffers may go out of scope after this
  if (h->slab && kept > c.max_markers) return fail(EULER_E_INVALID, "too many markers in slab");
  c.n_source_cells = (size_t)tot[0];
  c.n_source_cells_global = (size_t)tot[1];
  if (c.n_source_cells > h->source_cap || !c.source_cells) {
    int arc = alloc_array(h, &c.source_cells, c.n_source_cells);
    if (arc) return arc;
    h->source_cap = c.n_source_cells;
  }
  if (c.n_source_cells) launch_source_rows_write(c, h->src_rows);

  launch_refresh_counts(c);                                  // sim_init, main.c:268
  launch_colorize(c);                                        // --rainbow, main.c:271-273
  if (h->slab && h->comm_ready) {
    int grc = agree_on_sources(h);
    if (grc) return grc;
    if ((grc = color_halos(h))) return grc;
    // halo rows of the initial classification (only the owned markers were binned)
    if (comm_halo(c, h->cm, c.count, 1, SLAB_HALO)) return fail(EULER_E_COMM, "%s", comm_last_error());
  }
  int lrc = check_launch("load_state");
  if (lrc) return lrc;
  CU(cudaStreamSynchronize(c.stream));
  return 0;
}

// ---- project(), main.c:709-806 ------------------------------------------------------

void enqueue_precon_apply(euler_gpu* h, bool init) {
  if (h->prm.precon == EULER_PRECON_REDBLACK) launch_rb_apply(h->c, init);
  else launch_ic0_apply(h->c, init);
}

void enqueue_iteration(euler_gpu* h) {
  launch_apply_a(h->c, true);
  launch_axpy(h->c, h->prm.tol);
  enqueue_precon_apply(h, false);
  launch_update_search(h->c);
}

// How many iterations to enqueue before the host looks at the stop flag.  Consecutive solves of a
// run need about the same number of iterations (at scale they all end at the cap, main.c:735), so
// the first poll of a solve waits for as many iterations as the previous solve took: one host
// round trip per solve instead of one per `pcg_check_every` iterations.  Later polls go by
// `every`.  Iterations enqueued past convergence return at once (the stop flag is on the device).
int first_poll_chunk(const euler_gpu* h, int remaining, int every, bool first) {
  int chunk = every;
  if (first && h->poll_hint > every) chunk = h->poll_hint;
  return chunk < remaining ? chunk : remaining;
}

int run_project(euler_gpu* h, float dt) {
  Ctx& c = h->c;
  launch_build_rhs(c, dt);
  launch_tile_flags(c);
  int rc = pull_scalars(h);
  if (rc) return rc;
  if (h->host_sc->marker_overflow)
    return fail(EULER_E_UNSUPPORTED, "reference marker mode: more than %zu rewinding markers in one "
                "sub-step; use EULER_MARKERS_FAST", c.cand_cap);
  h->last_iterations = 0;
  if (!h->host_sc->nonzero_rhs) {
    h->solves_skipped++;                                    // all_zero(r), main.c:742
  } else {
    h->solves++;
    launch_pcg_reset(c);
    if (h->prm.precon == EULER_PRECON_REDBLACK) launch_rb_build(c);
    else launch_ic0_build(c);
    enqueue_precon_apply(h, true);                          // z = M^-1 r, sigma = z.r  (:744-748)
    bool first = true;
    if (!c.fused) launch_copy_search(c);                    // s = z                    (:746)
    int remaining = h->prm.max_iterations;
    const int every = h->prm.pcg_check_every > 0 ? h->prm.pcg_check_every : 8;
    bool first_chunk = true;
    // where iteration 1 leaves its s
    const void* s_odd = c.mixed ? static_cast<const void*>(c.s32b) : static_cast<const void*>(c.s2);
    const int refresh = c.mixed ? h->prm.pcg_refresh_every : 0;
    int it = 0;
    while (remaining > 0) {
      const int chunk = first_poll_chunk(h, remaining, every, first_chunk);
      first_chunk = false;
      for (int i = 0; i < chunk; ++i) {
        ++it;
        if (c.fused) {
          // c.z holds M^-1 r here; the fused kernels read it, leave A s in c.q ... and swap
          launch_fused_search_apply(c, first);              // s = z (+ beta s), A s -> c.q, alpha
          if (c.fused >= 2) {
            launch_fused_axpy_forward(c, h->prm.tol);       // p, r, ||r||inf, q = L^-1 r
          } else if (h->use_tail) {
            // r, (p), ||r||inf, q = L^-1 r, z = L^-T q, z.r, beta: one kernel (pcg_tail.cuh)
            launch_fused_tail(c, h->prm.tol, (it & 1) ? 0 : 1);
            first = false;
            continue;
          } else {
            launch_axpy(c, h->prm.tol, true, (it & 1) ? 0 : 1);  // r, ||r||inf; p every 2nd iteration
            // mixed precision: r <- b - A p in fp64 (p is complete after an even iteration)
            if (refresh > 0 && it % refresh == 0) launch_true_residual(c);
            launch_rb_forward(c);                           // q = L^-1 r -> c.q
          }
          launch_rb_backward(c, false);                     // z = L^-T q -> c.z, z.r, beta
          first = false;
        } else {
          enqueue_iteration(h);
        }
      }
      remaining -= chunk;
      rc = pull_scalars(h);
      if (rc) return rc;
      if (h->host_sc->done) break;
    }
    if (c.fused == 1) launch_p_fixup(c, s_odd);             // pending p update of an odd last iteration
    h->last_iterations = h->host_sc->iters;
    h->last_residual = h->host_sc->resid;
    h->pcg_iterations += (uint64_t)h->host_sc->iters;
    h->poll_hint = h->host_sc->iters;
  }
  launch_pressure_update(c, dt);
  h->max_valid = true;
  return check_launch("project");
}

int run_substep(euler_gpu* h, float dt) {
  Ctx& c = h->c;
  h->last_dt = dt;
  prof_mark(h, 0);
  launch_advect_markers(c, dt, h->prm.marker_mode);          // main.c:855
  launch_refresh_counts(c);                                  // main.c:856
  launch_extrapolate_color(c);                               // main.c:859-863 (--rainbow)
  launch_sources(c);                                         // main.c:864
  launch_source_colors(c, (unsigned int)h->frames);          // main.c:283, 292-294 (--rainbow)
  launch_grid_tiles(c);                                      // which tiles the grid stages stream
  prof_mark(h, 1);
  launch_extrapolate(c);                                     // main.c:865-868
  { float* t = c.u; c.u = c.uext; c.uext = t; t = c.v; c.v = c.vext; c.vext = t; }
  launch_advect_velocity(c, dt);                             // main.c:871-889
  launch_advect_color(c, dt);                                // main.c:873-882 (--rainbow)
  prof_mark(h, 2);
  int rc = run_project(h, dt);                               // main.c:893
  if (rc) return rc;
  prof_mark(h, 3);
  h->substeps++;
  c.gt_prev_sparse = c.gt_sparse;
  if (h->profiling) {
    CU(cudaEventSynchronize(h->ev[3]));
    CU(cudaStreamSynchronize(c.stream));
    prof_collect(c);
    float a = 0, b = 0, d = 0;
    cudaEventElapsedTime(&a, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&b, h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&d, h->ev[2], h->ev[3]);
    h->ms_markers += a; h->ms_grid += b; h->ms_project += d;
  }
  return 0;
}

// ---- row-slab mode: same stages, plus the exchange steps ---------------------
#define CM(call) do { if ((call)) return fail(EULER_E_COMM, "%s: %s", #call, comm_last_error()); } while (0)

int dist_after_backward(euler_gpu* h, bool init);

int dist_precon_apply(euler_gpu* h, bool init) {
  Ctx& c = h->c;
  // no exchange: r is kept consistent on the halo rows by redundant updates (pcg_kernels.cu
  // pview), q and z are recomputed there
  launch_rb_forward(c);
  launch_rb_backward(c, init);
  return dist_after_backward(h, init);
}

// the exchanges that follow z = M^-1 r on a slab: halo rows of z, {z.r, ||r||inf} across ranks
int dist_after_backward(euler_gpu* h, bool init) {
  Ctx& c = h->c;
  if (c.p2p_mode == 2) return 0;   // NVLink: halo stores + scalars happened inside k_rb_backward_pipe
  if (c.p2p_mode == 1) {
    // the same exchanges as two separate small kernels (A/B: EULER_P2P_SEPARATE=1)
    p2p_halo_z(c, h->cm, h->pp, SLAB_HALO);
    p2p_scalars(c, h->cm, h->pp, 1, init, h->prm.tol, true);
    return 0;
  }
  // one fused NCCL launch: the {z.r, ||r||inf} all-gather and — fused path — the halo exchange
  // of the new z that the next iteration's search/apply kernel starts from
  CM(comm_group_begin());
  CM(comm_gather_scalars(c, h->cm, c.sc->part));
  if (c.fused) {
    if (c.mixed) CM(comm_halo(c, h->cm, c.z32, 4, SLAB_HALO));
    else CM(comm_halo(c, h->cm, c.z, 8, SLAB_HALO));
  }
  CM(comm_group_end());
  launch_dist_beta(c, h->cm.gather, h->cm.nranks, init, h->prm.tol);
  return 0;
}

int dist_iteration(euler_gpu* h, bool first, int it) {
  Ctx& c = h->c;
  if (c.fused) {
    // the one exchange per iteration: z = M^-1 r, 4 rows deep; s' = z + beta s is then formed
    // redundantly on the halo rows (s itself was formed the same way one iteration earlier)
    // (z was exchanged together with the scalars at the end of the previous preconditioner)
    launch_fused_search_apply(c, first, h->split ? it : 0);   // s', A s' -> c.q, z.s partial
  } else {
    CM(comm_halo(c, h->cm, c.s, 8, SLAB_HALO));
    launch_apply_a(c, true);
  }
  if (c.p2p_mode == 2) {
    // {z.s} went over NVLink in the kernel's last block -> alpha
  } else if (c.p2p_mode == 1) {
    p2p_scalars(c, h->cm, h->pp, 0, false, h->prm.tol, false);
  } else {
    CM(comm_gather_scalars(c, h->cm, c.sc->part));  // {z.s partial}
    launch_dist_alpha(c, h->cm.gather, h->cm.nranks);
  }
  if (h->use_tail) {
    launch_fused_tail(c, h->prm.tol, (it & 1) ? 0 : 1, h->split ? it : 0);
    int trc = dist_after_backward(h, false);
    if (trc) return trc;
    return 0;
  }
  launch_axpy(c, h->prm.tol, c.fused != 0, c.fused == 1 ? ((it & 1) ? 0 : 1) : 2);
  // mixed precision: r <- b - A p in fp64 (p is complete after an even iteration, and kept current
  // on the +-3 halo rows by the redundant updates there)
  if (c.mixed && h->prm.pcg_refresh_every > 0 && it % h->prm.pcg_refresh_every == 0) launch_true_residual(c);
  int rc = dist_precon_apply(h, false);
  if (rc) return rc;
  if (!c.fused) launch_update_search(c);
  return 0;
}

int run_project_dist(euler_gpu* h, float dt) {
  Ctx& c = h->c;
  launch_build_rhs(c, dt);
  launch_tile_flags(c);
  CM(comm_allreduce_max_i32(c, h->cm, &c.sc->nonzero_rhs, 1));   // all_zero(r) over the whole grid
  int rc = pull_scalars(h);
  if (rc) return rc;
  h->last_iterations = 0;
  if (!h->host_sc->nonzero_rhs) {
    h->solves_skipped++;
  } else {
    h->solves++;
    launch_pcg_reset(c);
    launch_rb_build(c);
    CM(comm_halo(c, h->cm, c.r, 8, SLAB_HALO));      // b is only valid one row into the halo
    if (c.mixed) CM(comm_halo(c, h->cm, c.r32, 4, SLAB_HALO));
    rc = dist_precon_apply(h, true);
    if (rc) return rc;
    if (!c.fused) launch_copy_search(c);
    bool first = true;
    int remaining = h->prm.max_iterations;
    const int every = h->prm.pcg_check_every > 0 ? h->prm.pcg_check_every : 8;
    // where iteration 1 leaves its s
    const void* s_odd = c.mixed ? static_cast<const void*>(c.s32b) : static_cast<const void*>(c.s2);
    int it = 0;
    bool first_chunk = true;
    while (remaining > 0) {
      const int chunk = first_poll_chunk(h, remaining, every, first_chunk);
      first_chunk = false;
      for (int i = 0; i < chunk; ++i) { rc = dist_iteration(h, first, ++it); if (rc) return rc; first = false; }
      remaining -= chunk;
      // split-phase exchange: the last tail kernel's {z.r, ||r||inf} has no consumer yet — a
      // one-block kernel reads it for the host, and applies it when this was the last batch
      if (h->split) launch_dist_peek(c, remaining == 0);
      rc = pull_scalars(h);
      if (rc) return rc;
      if (h->host_sc->comm_timeout)
        return fail(EULER_E_COMM, "peer-to-peer exchange timed out (a rank stopped participating)");
      if (h->host_sc->done) break;
      if (h->split && h->host_sc->peek_done) {
        launch_dist_peek(c, true);
        rc = pull_scalars(h);
        if (rc) return rc;
        break;
      }
    }
    if (c.fused == 1) launch_p_fixup(c, s_odd, h->split ? 1 : 0);   // pending p update of an odd last iteration
    h->last_iterations = h->host_sc->iters;
    h->last_residual = h->host_sc->resid;
    h->pcg_iterations += (uint64_t)h->host_sc->iters;
    h->poll_hint = h->host_sc->iters;
  }
  // p[y+1] of the last owned row (main.c:800) is already there: p was updated redundantly on
  // the halo rows.  (When the solve was skipped p == 0 everywhere.)
  launch_pressure_update(c, dt);
  CM(comm_halo(c, h->cm, c.u, 4, SLAB_HALO));
  CM(comm_halo(c, h->cm, c.v, 4, SLAB_HALO));
  CM(comm_allreduce_max_u32(c, h->cm, &c.sc->max_u2_bits, 2));   // next calculate_timestep
  h->max_valid = true;
  return check_launch("project");
}

// markers that crossed into a neighbouring slab (displacement < 1 cell per sub-step): the advection
// kernel has put them into the staging buffers and turned the local copies into sink-column
// positions that the next refresh_marker_counts deletes; what the neighbours hand over is appended
int migrate_markers(euler_gpu* h) {
  Ctx& c = h->c;
  Comm& cm = h->cm;
  // counts first: what I send down/up, what the neighbours send me
  CU(cudaMemsetAsync(cm.mig, 0, 2 * sizeof(unsigned long long), c.stream));
  CM(comm_exchange(c, cm, &c.sc->n_send_dn, 8, &cm.mig[0], 8, &c.sc->n_send_up, 8, &cm.mig[1], 8));
  unsigned long long host[5];
  CU(cudaMemcpyAsync(&host[0], &c.sc->n_send_dn, 16, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaMemcpyAsync(&host[2], cm.mig, 16, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaMemcpyAsync(&host[4], &c.sc->n_markers, 8, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  const unsigned long long to_dn = host[0], to_up = host[1], from_dn = host[2], from_up = host[3], have = host[4];
  h->markers_migrated += to_dn + to_up;
  if (to_dn > cm.send_cap || to_up > cm.send_cap)
    return fail(EULER_E_UNSUPPORTED, "more than %zu markers cross a slab boundary in one sub-step", cm.send_cap);
  if (have + from_dn + from_up > c.max_markers)
    return fail(EULER_E_UNSUPPORTED, "slab marker capacity exceeded");
  CM(comm_exchange(c, cm, cm.send_dn, (size_t)to_dn * 8, c.markers + have, (size_t)from_dn * 8,
                   cm.send_up, (size_t)to_up * 8, c.markers + have + from_dn, (size_t)from_up * 8));
  if (from_dn + from_up) launch_add_markers(c, from_dn + from_up);
  return 0;
}

int run_substep_dist(euler_gpu* h, float dt) {
  Ctx& c = h->c;
  h->last_dt = dt;
  prof_mark(h, 0);
  launch_advect_markers_slab(c, dt, h->row0, h->row0 + h->rows, h->cm.rank > 0 ? h->cm.send_dn : nullptr,
                             h->cm.rank + 1 < h->cm.nranks ? h->cm.send_up : nullptr, h->cm.send_cap);
  int rc = migrate_markers(h);
  if (rc) return rc;
  launch_refresh_counts(c);
  if (c.n_source_cells_global) {
    launch_sources_count(c);
    CM(comm_gather_scalars(c, h->cm, c.sc->part));
    launch_sources_prep(c, h->cm.gather, h->cm.rank, h->cm.nranks);
    launch_sources(c);
  }
  CM(comm_halo(c, h->cm, c.count, 1, SLAB_HALO));    // classification of the neighbours' edge rows
  // --rainbow: extrapolate(P) (main.c:859-863) needs the neighbours' classification, so it runs
  // after the halo exchange instead of before the sources; it reads only colours of cells that were
  // fluid LAST sub-step and the sources' colour writes (main.c:292-294) still come after it, so the
  // result is the reference's
  launch_extrapolate_color(c);
  launch_source_colors(c, (unsigned int)h->frames);
  launch_grid_tiles(c);
  prof_mark(h, 1);
  launch_extrapolate(c);
  { float* t = c.u; c.u = c.uext; c.uext = t; t = c.v; c.v = c.vext; c.vext = t; }
  launch_advect_velocity(c, dt);
  launch_advect_color(c, dt);                        // main.c:873-882
  { int grc = color_halos(h); if (grc) return grc; }
  prof_mark(h, 2);
  rc = run_project_dist(h, dt);
  if (rc) return rc;
  prof_mark(h, 3);
  h->substeps++;
  c.gt_prev_sparse = c.gt_sparse;
  if (h->profiling) {
    CU(cudaEventSynchronize(h->ev[3]));
    CU(cudaStreamSynchronize(c.stream));
    prof_collect(c);
    float a = 0, b = 0, d = 0;
    cudaEventElapsedTime(&a, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&b, h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&d, h->ev[2], h->ev[3]);
    h->ms_markers += a; h->ms_grid += b; h->ms_project += d;
  }
  return 0;
}

int compute_dt(euler_gpu* h, float frame_time, float* dt) {
  if (!h->max_valid) {
    launch_maxsq(h->c);
    if (h->slab) CM(comm_allreduce_max_u32(h->c, h->cm, &h->c.sc->max_u2_bits, 2));
  }
  h->max_valid = true;
  launch_timestep(h->c, frame_time, h->prm.cfl_distance);
  int rc = pull_scalars(h);
  if (rc) return rc;
  *dt = h->host_sc->dt;
  return 0;
}

}  // namespace

// ======================================================================= C-ABI ====

extern "C" {

int euler_gpu_abi_version(void) { return EULER_GPU_ABI_VERSION; }
const char* euler_gpu_last_error(void) { return g_err; }

int euler_gpu_default_params(euler_params* p) {
  if (!p) return fail(EULER_E_INVALID, "params is NULL");
  memset(p, 0, sizeof *p);
  p->h = 1.f; p->rho = 1.f; p->gravity = -10.f;              // main.c:58-60
  p->frame_time = 0.1f; p->max_substeps = 8;                 // main.c:849-851
  p->cfl_distance = 0.75f;                                   // main.c:838
  p->max_iterations = 100;                                   // main.c:735
  p->tol = (double)1e-6f;                                    // main.c:736
  p->precon = EULER_PRECON_IC0_WAVEFRONT;
  p->marker_mode = EULER_MARKERS_REFERENCE;
  p->dot_mode = EULER_DOT_TREE;
  p->rng_state = 0x9bd185c449534b91ull;                      // main.c:204
  p->device = 0; p->stream = nullptr; p->pcg_check_every = 8;
  p->slab_row0 = 0; p->slab_rows = 0;
  p->pcg_dtype = EULER_PCG_FP64; p->pcg_refresh_every = 10;
  return 0;
}

int euler_gpu_destroy(euler_gpu* h) {
  if (!h) return 0;
  cudaSetDevice(h->prm.device);
  if (h->c.stream) cudaStreamSynchronize(h->c.stream);
  p2p_close(h->pp, h->cm);
  if (h->comm_ready) comm_destroy(&h->cm);
  for (void* p : h->allocs) cudaFree(p);
  if (h->host_sc) cudaFreeHost(h->host_sc);
  for (int i = 0; i < 4; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->c.prof.ev) {
    for (int i = 0; i < 2 * h->c.prof.cap; ++i) if (h->c.prof.ev[i]) cudaEventDestroy(h->c.prof.ev[i]);
    delete[] h->c.prof.ev; delete[] h->c.prof.cls;
  }
  if (h->up_stream) { cudaStreamDestroy(h->up_stream); cudaEventDestroy(h->up_event); }
  if (h->own_stream && h->c.stream) cudaStreamDestroy(h->c.stream);
  delete h;
  return 0;
}

int euler_gpu_create(euler_gpu** out, int nx, int ny, const uint8_t* solid, const uint8_t* source,
                     const uint8_t* sink, const float* markers_xy, size_t n_markers,
                     const euler_params* params) {
  if (!out) return fail(EULER_E_INVALID, "out is NULL");
  *out = nullptr;
  if (nx < 4 || ny < 4) return fail(EULER_E_INVALID, "grid %dx%d too small (min 4x4)", nx, ny);
  if (!solid || !source || !sink) return fail(EULER_E_INVALID, "mask planes must not be NULL");
  if (n_markers && !markers_xy) return fail(EULER_E_INVALID, "markers is NULL");
  const size_t max_markers_global = 4 * (size_t)nx * ny;     // MAX_MARKER_COUNT, main.c:92
  if (n_markers > max_markers_global) return fail(EULER_E_INVALID, "too many markers");
  if (max_markers_global >= 0xFFFFFFFFull) return fail(EULER_E_UNSUPPORTED, "grid too large");
  euler_params prm;
  if (params) prm = *params; else euler_gpu_default_params(&prm);
  const bool slab = prm.slab_rows > 0;
  if (slab) {
    if (prm.slab_row0 < 0 || prm.slab_row0 + prm.slab_rows > ny || prm.slab_rows < 2 * SLAB_HALO)
      return fail(EULER_E_INVALID, "slab rows [%d,%d) invalid for ny=%d (min %d rows per slab)",
                  prm.slab_row0, prm.slab_row0 + prm.slab_rows, ny, 2 * SLAB_HALO);
    if (prm.precon != EULER_PRECON_REDBLACK || prm.marker_mode != EULER_MARKERS_FAST || prm.dot_mode != EULER_DOT_TREE)
      return fail(EULER_E_UNSUPPORTED, "row slabs need precon=REDBLACK, marker_mode=FAST, dot_mode=TREE "
                  "(the IC(0) wavefront and the reference orders are sequential across the whole grid)");
  }
  if (prm.pcg_dtype != EULER_PCG_FP64 && prm.pcg_dtype != EULER_PCG_FP32)
    return fail(EULER_E_INVALID, "unknown pcg_dtype %d", prm.pcg_dtype);
  const bool mixed = prm.pcg_dtype == EULER_PCG_FP32;
  if (mixed) {
    if (prm.precon != EULER_PRECON_REDBLACK || prm.dot_mode != EULER_DOT_TREE || prm.stencil_variant != 0)
      return fail(EULER_E_UNSUPPORTED, "pcg_dtype=FP32 needs precon=REDBLACK, dot_mode=TREE, stencil_variant=0 "
                  "(it is a mode of the fused red-black iteration)");
    if (prm.pcg_refresh_every < 0 || (prm.pcg_refresh_every & 1))
      return fail(EULER_E_INVALID, "pcg_refresh_every must be even (p is complete after even iterations) or 0, got %d",
                  prm.pcg_refresh_every);
  }
  const int row0 = slab ? prm.slab_row0 : 0, rows = slab ? prm.slab_rows : ny;
  const int lo = row0 - SLAB_HALO > 0 ? row0 - SLAB_HALO : 0;
  const int hi = row0 + rows + SLAB_HALO < ny ? row0 + rows + SLAB_HALO : ny;
  const int ny_loc = slab ? hi - lo : ny;
  // local capacity: every stored row full, plus what may arrive from the neighbours
  const size_t max_markers = 4 * (size_t)nx * ny_loc;
  if (prm.precon != EULER_PRECON_IC0_WAVEFRONT && prm.precon != EULER_PRECON_REDBLACK)
    return fail(EULER_E_INVALID, "unknown preconditioner %d", prm.precon);
  if (prm.marker_mode != EULER_MARKERS_REFERENCE && prm.marker_mode != EULER_MARKERS_FAST)
    return fail(EULER_E_INVALID, "unknown marker mode %d", prm.marker_mode);

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(EULER_E_CUDA, "no CUDA device: %s (libeuler_gpu has no CPU fallback)",
                cudaGetErrorString(e));
  if (prm.device < 0 || prm.device >= ndev) return fail(EULER_E_INVALID, "device %d of %d", prm.device, ndev);
  CU(cudaSetDevice(prm.device));

  euler_gpu* h = new euler_gpu();
  memset(&h->c, 0, sizeof h->c);
  h->prm = prm; h->nx = nx; h->ny = ny;
  h->slab = slab; h->row0 = row0; h->rows = rows; h->lo = slab ? lo : 0;
  memset(&h->cm, 0, sizeof h->cm); h->comm_ready = false; h->n_keep = nullptr;
  memset(&h->pp, 0, sizeof h->pp); h->z_raw = nullptr; h->source_cap = 0;
  h->host_sc = nullptr; h->device_bytes = 0; h->max_valid = false;
  h->frames = h->substeps = h->solves = h->solves_skipped = h->pcg_iterations = h->markers_migrated = 0;
  h->last_iterations = 0; h->last_residual = 0; h->last_dt = 0; h->profiling = false; h->poll_hint = 0;
  h->up_stream = nullptr; h->up_event = nullptr; h->src_rows = nullptr; h->src_totals = nullptr;
  h->split = false;
  h->ms_markers = h->ms_grid = h->ms_project = 0;
  for (int i = 0; i < 4; ++i) h->ev[i] = nullptr;
  Ctx& c = h->c;
  c.g.nx = nx; c.g.ny = ny_loc;
  c.g.pitch = (nx + PITCH_ALIGN - 1) / PITCH_ALIGN * PITCH_ALIGN;
  c.g.yoff = h->lo; c.g.gny = ny; c.g.th = 32;
  c.own0 = row0 - h->lo; c.own1 = c.own0 + rows;
  c.distributed = 0;
  c.max_markers_global = max_markers_global;
  c.lim.u_x = nextafterf((float)(nx - 2), 0.f);              // main.c:339-340, U is (X-1) x Y
  c.lim.u_y = nextafterf((float)(ny - 1), 0.f);
  c.lim.v_x = nextafterf((float)(nx - 1), 0.f);              // V is X x (Y-1)
  c.lim.v_y = nextafterf((float)(ny - 2), 0.f);
  c.lim.p_x = nextafterf((float)(nx - 1), 0.f);              // P is X x Y
  c.lim.p_y = nextafterf((float)(ny - 1), 0.f);
  c.h = prm.h; c.rho = prm.rho; c.gravity = prm.gravity;
  c.dot_mode = prm.dot_mode ? 1 : 0;
  c.tol = prm.tol;
  c.use_pipe = prm.stencil_variant != 1 ? 1 : 0;
  c.max_markers = max_markers;

#define TRY(x) do { int rc_ = (x); if (rc_) { euler_gpu_destroy(h); return rc_; } } while (0)
#define TRYCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { int rc_ = fail(EULER_E_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); euler_gpu_destroy(h); return rc_; } } while (0)

  cudaDeviceProp prop;
  TRYCU(cudaGetDeviceProperties(&prop, prm.device));
  c.sm_count = prop.multiProcessorCount;
  if (prm.stream) { c.stream = (cudaStream_t)prm.stream; h->own_stream = false; }
  else { TRYCU(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking)); h->own_stream = true; }
  for (int i = 0; i < 4; ++i) TRYCU(cudaEventCreate(&h->ev[i]));
  c.prof.cap = 4096;
  c.prof.ev = new cudaEvent_t[2 * c.prof.cap]();
  c.prof.cls = new int[c.prof.cap]();
  TRYCU(cudaMallocHost((void**)&h->host_sc, sizeof(DevScalars)));

  TRY(alloc_plane(h, &c.solid)); TRY(alloc_plane(h, &c.source)); TRY(alloc_plane(h, &c.sink));
  TRY(alloc_plane(h, &c.count)); TRY(alloc_plane(h, &c.prev_count));
  TRY(alloc_plane(h, &c.count32));
  TRY(alloc_plane(h, &c.u)); TRY(alloc_plane(h, &c.v));
  TRY(alloc_plane(h, &c.utmp)); TRY(alloc_plane(h, &c.vtmp));
  TRY(alloc_plane(h, &c.uext)); TRY(alloc_plane(h, &c.vext));
  TRY(alloc_plane(h, &c.adiag));
  // stencil_variant 0: fused update_search+apply_a (default); 2: additionally axpy+forward fused
  // (measured slower: profiles/r01 notes); 1: nothing fused, register-window stencils
  c.fused = (prm.precon == EULER_PRECON_REDBLACK && prm.dot_mode == EULER_DOT_TREE && prm.stencil_variant != 1)
                ? (prm.stencil_variant == 2 ? 2 : 1) : 0;
  c.mixed = mixed ? 1 : 0;
  {
    // EULER_TAIL=0: the three separate kernels (k_axpy, k_rb_forward_pipe, k_rb_backward_pipe) for A/B runs
    const char* te = getenv("EULER_TAIL");
    h->use_tail = c.fused == 1 && !mixed && !(te && atoi(te) == 0);
  }
  {
    // ring depths of the fp32 pipe kernels (A/B knobs: EULER_NS_MIXED for all three, or
    // EULER_NS_MIXED_F / _B / _KA), read once per handle
    static const char* const names[3] = {"EULER_NS_MIXED_F", "EULER_NS_MIXED_B", "EULER_NS_MIXED_KA"};
    const char* all = getenv("EULER_NS_MIXED");
    for (int i = 0; i < 3; ++i) {
      const char* e = getenv(names[i]);
      const int v = e ? atoi(e) : all ? atoi(all) : NS_MIXED_DEFAULT[i];
      c.ns_mixed[i] = (v == 6 || v == 8) ? v : 4;
    }
    // EULER_MIXED_BLOCKS=8: the fp32 pipe kernels compiled for 8 resident blocks per SM (64
    // registers; the backward solve spills 16 B) instead of 7 — unmeasured, off by default
    const char* mb = getenv("EULER_MIXED_BLOCKS");
    c.mixed_blocks = (mb && atoi(mb) == 8) ? 8 : 0;
  }
  if (mixed) {
    // fp64: p and b (in the r plane) only; everything the iteration streams is fp32
    TRY(alloc_plane(h, &c.p)); TRY(alloc_plane(h, &c.r));
    TRY(alloc_plane(h, &c.r32)); TRY(alloc_plane(h, &c.z32)); h->z_raw = h->allocs.back();
    TRY(alloc_plane(h, &c.s32));
    TRY(alloc_plane(h, &c.s32b)); TRY(alloc_plane(h, &c.q32)); TRY(alloc_plane(h, &c.pc32));
  } else {
    TRY(alloc_plane(h, &c.precon)); TRY(alloc_plane(h, &c.q)); TRY(alloc_plane(h, &c.p));
    TRY(alloc_plane(h, &c.r)); TRY(alloc_plane(h, &c.z)); h->z_raw = h->allocs.back();
    TRY(alloc_plane(h, &c.s));
    if (c.fused) { TRY(alloc_plane(h, &c.s2)); TRY(alloc_plane(h, &c.r2)); }
  }
  if (prm.rainbow) {
    TRY(alloc_plane(h, &c.cr)); TRY(alloc_plane(h, &c.cg)); TRY(alloc_plane(h, &c.cb));
    TRY(alloc_plane(h, &c.crtmp)); TRY(alloc_plane(h, &c.cgtmp)); TRY(alloc_plane(h, &c.cbtmp));
  }
  TRY(alloc_array(h, &c.markers, max_markers));
  TRY(alloc_array(h, &c.markers_alt, max_markers));
  c.n_segments = (max_markers + 1023) / 1024;
  TRY(alloc_array(h, &c.seg_count, c.n_segments));
  TRY(alloc_array(h, &c.seg_offset, c.n_segments));
  if (prm.marker_mode == EULER_MARKERS_REFERENCE) {
    // rewinding markers are rare (they crossed a cell edge and then hit a wall in the same
    // sub-step); 1/16 of the marker capacity, at least 64k records
    c.cand_cap = max_markers / 16 > 65536 ? max_markers / 16 : 65536;
    unsigned char* raw = nullptr;
    TRY(alloc_array(h, &raw, c.cand_cap * marker_candidate_bytes()));
    c.cand = raw;
    TRY(alloc_array(h, &c.cand_dt, c.cand_cap));
  }
  const size_t nblk2d = (size_t)((nx + 31) / 32) * (size_t)((ny + 7) / 8);
  c.n_strips = (ny_loc - 2 + 31) / 32;
  c.n_partials = 65536 > (size_t)c.n_strips ? 65536 : (size_t)c.n_strips;
  (void)nblk2d;
  c.gt_tx = (c.g.pitch + GT_W - 1) / GT_W; c.gt_ty = (c.g.ny + GT_H - 1) / GT_H;
  for (int i = 0; i < 3; ++i) TRY(alloc_array(h, &c.gt_flags[i], (size_t)c.gt_tx * c.gt_ty));
  TRY(alloc_array(h, &c.gt_list, (size_t)c.gt_tx * c.gt_ty));
  c.gt_hist = c.gt_sparse = c.gt_prev_sparse = 0;
  TRY(alloc_array(h, &c.tile_active, (size_t)pcg_tile_count(c.g)));
  TRY(alloc_array(h, &c.tile_list, (size_t)pcg_tile_count(c.g)));
  TRY(alloc_array(h, &c.partials, c.n_partials));
  TRY(alloc_array(h, &c.wf_progress, (size_t)c.n_strips));
  TRY(alloc_array(h, &c.sc, 1));
  TRY(alloc_array(h, &c.rng_jump, 64 * 64));

  {
    std::vector<unsigned long long> jump(64 * 64);
    init_rng_jump_table(jump.data());
    TRYCU(cudaMemcpyAsync(c.rng_jump, jump.data(), jump.size() * 8, cudaMemcpyHostToDevice, c.stream));
    TRYCU(cudaStreamSynchronize(c.stream));
  }
  if (slab) TRY(alloc_array(h, &h->n_keep, 1));
  TRY(alloc_array(h, &h->src_rows, (size_t)c.g.ny + 1));
  TRY(alloc_array(h, &h->src_totals, 2));
  {
    // EULER_TRACE=<slots>: in-kernel timeline of the PCG iteration kernels (euler_gpu_trace_read)
    const char* te = getenv("EULER_TRACE");
    const int slots = te ? atoi(te) : 0;
    c.trace = nullptr; c.trace_cap = 0; c.trace_n = 0;
    if (slots > 0) {
      TRY(alloc_array(h, &c.trace, (size_t)slots * TRACE_WORDS + (size_t)TRACE_BLK_MAX * TRACE_BLK_WORDS));
      c.trace_cap = slots;
    }
  }
  // (slab handles: the halo rows of the count plane are filled by comm_init's halo exchange)
  TRY(load_state(h, solid, source, sink, markers_xy, n_markers, prm.rng_state, true));
#undef TRY
#undef TRYCU
  *out = h;
  return 0;
}

#define ENTER0(h)                                                  \
  if (!(h)) return fail(EULER_E_INVALID, "handle is NULL");        \
  CU(cudaSetDevice((h)->prm.device))

int euler_gpu_reinit(euler_gpu* h, const uint8_t* solid, const uint8_t* source, const uint8_t* sink,
                     const float* markers_xy, size_t n_markers, uint64_t rng_state) {
  ENTER0(h);
  if (!solid || !source || !sink) return fail(EULER_E_INVALID, "mask planes must not be NULL");
  if (n_markers && !markers_xy) return fail(EULER_E_INVALID, "markers is NULL");
  if (n_markers > h->c.max_markers_global) return fail(EULER_E_INVALID, "too many markers");
  if (h->slab && !h->comm_ready) return fail(EULER_E_COMM, "slab handle: call euler_gpu_comm_init first");
  CU(cudaStreamSynchronize(h->c.stream));
  h->prm.rng_state = rng_state;
  h->frames = 0;                 // g_frame_count (main.c:89) is simulation state: the source hue restarts at t = 0
  return load_state(h, solid, source, sink, markers_xy, n_markers, rng_state, false);
}

#define ENTER(h)                                                   \
  if (!(h)) return fail(EULER_E_INVALID, "handle is NULL");        \
  CU(cudaSetDevice((h)->prm.device))
#define NEED_COMM(h) \
  if ((h)->slab && !(h)->comm_ready) return fail(EULER_E_COMM, "slab handle: call euler_gpu_comm_init first")

int euler_gpu_calculate_timestep(euler_gpu* h, float frame_time, float* dt) {
  ENTER(h);
  NEED_COMM(h);
  if (!dt) return fail(EULER_E_INVALID, "dt is NULL");
  return compute_dt(h, frame_time, dt);
}

int euler_gpu_substep(euler_gpu* h, float dt) {
  ENTER(h);
  NEED_COMM(h);
  int rc = h->slab ? run_substep_dist(h, dt) : run_substep(h, dt);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->c.stream));
  return 0;
}

int euler_gpu_step_frame(euler_gpu* h, int* substeps) {
  ENTER(h);
  NEED_COMM(h);
  float frame_time = h->prm.frame_time;                      // main.c:849-851
  int step = 0;
  for (; frame_time > 0.f && step < h->prm.max_substeps; ++step) {
    float dt = 0.f;
    int rc = compute_dt(h, frame_time, &dt);
    if (rc) return rc;
    frame_time -= dt;
    rc = h->slab ? run_substep_dist(h, dt) : run_substep(h, dt);
    if (rc) return rc;
  }
  h->frames++;
  if (substeps) *substeps = step;
  CU(cudaStreamSynchronize(h->c.stream));
  return check_launch("step_frame");
}

int euler_gpu_run_stage(euler_gpu* h, int stage, float dt) {
  ENTER(h);
  if (h->slab) return fail(EULER_E_UNSUPPORTED, "run_stage is a single-GPU parity hook");
  Ctx& c = h->c;
  c.gt_hist = c.gt_sparse = c.gt_prev_sparse = 0;            // single stages always stream every tile
  switch (stage) {
    case EULER_S_ADVECT_MARKERS: launch_advect_markers(c, dt, h->prm.marker_mode); break;
    case EULER_S_REFRESH_COUNTS: launch_refresh_counts(c); break;
    case EULER_S_SOURCES: launch_sources(c); break;
    case EULER_S_EXTRAPOLATE: {
      launch_extrapolate(c);
      float* t = c.u; c.u = c.uext; c.uext = t; t = c.v; c.v = c.vext; c.vext = t;
      h->max_valid = false;
      break;
    }
    case EULER_S_ADVECT_VELOCITY: launch_advect_velocity(c, dt); break;
    case EULER_S_PROJECT: { int rc = run_project(h, dt); if (rc) return rc; break; }
    case EULER_S_BUILD_RHS: launch_build_rhs(c, dt); launch_tile_flags(c); break;
    case EULER_S_PRECONDITION:
      launch_pcg_reset(c);
      launch_tile_flags(c);
      if (h->prm.precon == EULER_PRECON_REDBLACK) launch_rb_build(c); else launch_ic0_build(c);
      enqueue_precon_apply(h, true);
      break;
    case EULER_S_APPLY_A:
      if (c.mixed) return fail(EULER_E_UNSUPPORTED, "pcg_dtype=FP32 has no stand-alone apply_a (fused into the search update)");
      launch_pcg_reset(c); launch_tile_flags(c); launch_apply_a(c, false); break;
    case EULER_S_PRESSURE_UPDATE: launch_pressure_update(c, dt); h->max_valid = true; break;
    case EULER_S_FUSED_TAIL:
      if (c.fused != 1 || c.mixed)
        return fail(EULER_E_UNSUPPORTED, "the fused tail is part of the fp64 fused red-black iteration (precon=REDBLACK, dot_mode=TREE, stencil_variant=0)");
      launch_pcg_reset(c); launch_tile_flags(c); launch_rb_build(c);
      launch_set_alpha(c, (double)dt);
      launch_fused_tail(c, h->prm.tol, 2);
      break;
    case EULER_S_EXTRAPOLATE_COLOR:
    case EULER_S_ADVECT_COLOR:
      if (!c.cr) return fail(EULER_E_INVALID, "handle was created without params.rainbow");
      if (stage == EULER_S_EXTRAPOLATE_COLOR) launch_extrapolate_color(c); else launch_advect_color(c, dt);
      break;
    default: return fail(EULER_E_INVALID, "unknown stage %d", stage);
  }
  int rc = check_launch("run_stage");
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->c.stream));
  return check_launch("run_stage");
}

int euler_gpu_pcg_iterations(euler_gpu* h, int iterations) {
  ENTER(h);
  if (h->c.mixed) return fail(EULER_E_UNSUPPORTED, "pcg_iterations drives the unfused fp64 iteration");
  for (int i = 0; i < iterations; ++i) {
    enqueue_iteration(h);
    if (h->c.prof.on && h->c.prof.n + 16 > h->c.prof.cap) {
      CU(cudaStreamSynchronize(h->c.stream));
      prof_collect(h->c);
    }
  }
  return check_launch("pcg_iterations");
}

int euler_gpu_read_marker_count(euler_gpu* h, uint8_t* dst) {
  return euler_gpu_get(h, EULER_F_COUNT, dst, (size_t)(h ? h->nx : 0) * (h ? h->ny : 0));
}

int euler_gpu_get(euler_gpu* h, int field, void* dst, size_t bytes) {
  ENTER(h);
  if (!dst) return fail(EULER_E_INVALID, "dst is NULL");
  const Grid& g = h->c.g;
  if (field == EULER_F_MARKERS) {
    int rc = pull_scalars(h);
    if (rc) return rc;
    const size_t want = (size_t)h->host_sc->n_markers * sizeof(float2);
    if (bytes != want) return fail(EULER_E_INVALID, "markers: %zu bytes given, %zu needed", bytes, want);
    if (want) CU(cudaMemcpyAsync(dst, h->c.markers, want, cudaMemcpyDeviceToHost, h->c.stream));
    CU(cudaStreamSynchronize(h->c.stream));
    return 0;
  }
  FieldInfo fi = field_info(h, field);
  if (!fi.ptr) return fail(EULER_E_INVALID, "field %d unknown or not allocated on this handle (params.rainbow / pcg_dtype)", field);
  const size_t want = (size_t)g.nx * h->ny * fi.elem;
  if (bytes != want) return fail(EULER_E_INVALID, "field %d: %zu bytes given, %zu needed", field, bytes, want);
  // slab mode: only the rows this handle OWNS are written (at their global position)
  char* to = reinterpret_cast<char*>(dst) + (size_t)h->row0 * g.nx * fi.elem;
  const char* from = reinterpret_cast<const char*>(fi.ptr) + (size_t)h->c.own0 * g.pitch * fi.elem;
  CU(cudaMemcpy2DAsync(to, (size_t)g.nx * fi.elem, from, g.pitch * fi.elem, (size_t)g.nx * fi.elem,
                       h->rows, cudaMemcpyDeviceToHost, h->c.stream));
  CU(cudaStreamSynchronize(h->c.stream));
  return 0;
}

int euler_gpu_read_window(euler_gpu* h, int field, int x0, int y0, int w, int hh, void* dst_global) {
  ENTER(h);
  if (!dst_global) return fail(EULER_E_INVALID, "dst is NULL");
  FieldInfo fi = field_info(h, field);
  if (!fi.ptr) return fail(EULER_E_INVALID, "field %d is not a plane", field);
  if (x0 < 0 || y0 < 0 || w < 0 || hh < 0 || x0 + w > h->nx || y0 + hh > h->ny)
    return fail(EULER_E_INVALID, "window [%d,%d)x[%d,%d) outside the %dx%d grid", x0, x0 + w, y0, y0 + hh, h->nx, h->ny);
  const Grid& g = h->c.g;
  // slab mode: the part of the window inside the rows this handle owns
  const int ya = y0 > h->row0 ? y0 : h->row0;
  const int yb = y0 + hh < h->row0 + h->rows ? y0 + hh : h->row0 + h->rows;
  if (w > 0 && yb > ya) {
    char* to = reinterpret_cast<char*>(dst_global) + ((size_t)ya * g.nx + x0) * fi.elem;
    const char* from = reinterpret_cast<const char*>(fi.ptr) + ((size_t)(ya - h->lo) * g.pitch + x0) * fi.elem;
    CU(cudaMemcpy2DAsync(to, (size_t)g.nx * fi.elem, from, g.pitch * fi.elem, (size_t)w * fi.elem,
                         (size_t)(yb - ya), cudaMemcpyDeviceToHost, h->c.stream));
  }
  CU(cudaStreamSynchronize(h->c.stream));
  return 0;
}

int euler_gpu_set(euler_gpu* h, int field, const void* src, size_t bytes) {
  ENTER(h);
  if (!src && bytes) return fail(EULER_E_INVALID, "src is NULL");
  const Grid& g = h->c.g;
  // the caller may put values (or markers) where the grid stages' tile list assumes zeros: the
  // next sub-steps stream every tile again until the list's memory is rebuilt
  h->c.gt_hist = h->c.gt_sparse = h->c.gt_prev_sparse = 0;
  if (field == EULER_F_MARKERS) {
    if (bytes % sizeof(float2)) return fail(EULER_E_INVALID, "markers: size not a multiple of 8");
    const size_t n = bytes / sizeof(float2);
    if (h->slab) return fail(EULER_E_UNSUPPORTED, "set(MARKERS) on a slab handle");
    if (n > h->c.max_markers) return fail(EULER_E_INVALID, "markers: %zu > max %zu", n, h->c.max_markers);
    if (n) CU(cudaMemcpyAsync(h->c.markers, src, bytes, cudaMemcpyHostToDevice, h->c.stream));
    unsigned long long nn = n;
    CU(cudaMemcpyAsync(&h->c.sc->n_markers, &nn, sizeof nn, cudaMemcpyHostToDevice, h->c.stream));
    CU(cudaStreamSynchronize(h->c.stream));
    return 0;
  }
  if (field == EULER_F_SOURCE) return fail(EULER_E_UNSUPPORTED, "source plane is fixed at create()");
  FieldInfo fi = field_info(h, field);
  if (!fi.ptr) return fail(EULER_E_INVALID, "field %d unknown or not allocated on this handle (params.rainbow / pcg_dtype)", field);
  const size_t want = (size_t)g.nx * h->ny * fi.elem;
  if (bytes != want) return fail(EULER_E_INVALID, "field %d: %zu bytes given, %zu needed", field, bytes, want);
  int rc = upload_plane(h, fi.ptr, src, fi.elem, h->c.stream);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->c.stream));
  if (field == EULER_F_U || field == EULER_F_V) h->max_valid = false;
  return 0;
}

int euler_gpu_colorize(euler_gpu* h) {
  ENTER(h);
  if (!h->c.cr) return fail(EULER_E_INVALID, "handle was created without params.rainbow");
  launch_colorize(h->c);
  CU(cudaStreamSynchronize(h->c.stream));
  return check_launch("colorize");
}

int euler_gpu_set_rng_state(euler_gpu* h, uint64_t state) {
  ENTER(h);
  unsigned long long s = state;
  CU(cudaMemcpyAsync(&h->c.sc->rng_state, &s, sizeof s, cudaMemcpyHostToDevice, h->c.stream));
  CU(cudaStreamSynchronize(h->c.stream));
  return 0;
}

int euler_gpu_set_max_iterations(euler_gpu* h, int max_iterations) {
  ENTER(h);
  if (max_iterations < 0) return fail(EULER_E_INVALID, "max_iterations %d", max_iterations);
  h->prm.max_iterations = max_iterations;
  h->poll_hint = 0;
  return 0;
}

int euler_gpu_set_frame_count(euler_gpu* h, uint64_t frames) {
  ENTER(h);
  h->frames = frames;            // g_frame_count (main.c:89): only the source colours read it
  return 0;
}

int euler_gpu_set_source_exhausted(euler_gpu* h, int exhausted) {
  ENTER(h);
  int v = exhausted ? 1 : 0;
  CU(cudaMemcpyAsync(&h->c.sc->source_exhausted, &v, sizeof v, cudaMemcpyHostToDevice, h->c.stream));
  CU(cudaStreamSynchronize(h->c.stream));
  return 0;
}

int euler_gpu_stats(euler_gpu* h, euler_stats* out) {
  ENTER(h);
  if (!out) return fail(EULER_E_INVALID, "out is NULL");
  int rc = pull_scalars(h);
  if (rc) return rc;
  memset(out, 0, sizeof *out);
  out->frames = h->frames; out->substeps = h->substeps;
  out->solves = h->solves; out->solves_skipped = h->solves_skipped;
  out->pcg_iterations = h->pcg_iterations;
  out->last_iterations = h->last_iterations;
  out->last_residual = h->last_residual;
  out->last_dt = h->last_dt;
  out->n_markers = h->host_sc->n_markers;
  out->source_exhausted = h->host_sc->source_exhausted;
  out->rng_state = h->host_sc->rng_state;
  out->kernel_launches = h->c.launches;
  out->device_bytes = h->device_bytes;
  out->ms_markers = h->ms_markers; out->ms_grid = h->ms_grid; out->ms_project = h->ms_project;
  out->active_cells = (uint64_t)h->host_sc->active_tiles * (uint64_t)pcg_tile_cells(h->c);
  out->markers_migrated = h->markers_migrated;
  out->grid_cells = h->c.gt_prev_sparse ? (uint64_t)h->host_sc->grid_tiles * (uint64_t)(GT_W * GT_H)
                                        : (uint64_t)h->c.g.ny * (uint64_t)h->c.g.pitch;
  for (int i = 0; i < KC__COUNT; ++i) { out->kernel_ms[i] = h->c.prof.ms[i]; out->kernel_count[i] = h->c.prof.count[i]; }
  return 0;
}

int euler_gpu_check(euler_gpu* h, euler_check* out) {
  ENTER(h);
  if (!out) return fail(EULER_E_INVALID, "out is NULL");
  Ctx& c = h->c;
  const int blocks = 1024;                       // c.partials holds >= 65536 doubles
  unsigned long long* ints = reinterpret_cast<unsigned long long*>(c.partials + 6 * blocks);
  launch_check(c, ints, c.partials, blocks);
  std::vector<double> parts(6 * blocks);
  unsigned long long hi[3];
  CU(cudaMemcpyAsync(parts.data(), c.partials, parts.size() * 8, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaMemcpyAsync(hi, ints, sizeof hi, cudaMemcpyDeviceToHost, c.stream));
  int rc = pull_scalars(h);                      // synchronises the stream
  if (rc) return rc;
  memset(out, 0, sizeof *out);
  out->n_markers = h->host_sc->n_markers;
  out->fluid_cells = hi[0]; out->count_sum = hi[1]; out->count_hash = hi[2];
  for (int b = 0; b < blocks; ++b) {
    const double* q = &parts[6 * b];
    out->sum_abs_u += q[0]; out->sum_abs_v += q[1]; out->sum_p += q[2];
    if (q[3] > out->max_abs_div) out->max_abs_div = q[3];
    if (q[4] > out->max_abs_u) out->max_abs_u = q[4];
    if (q[5] > out->max_abs_v) out->max_abs_v = q[5];
  }
  return check_launch("check");
}

int euler_gpu_set_profiling(euler_gpu* h, int enabled) {
  ENTER(h);
  CU(cudaStreamSynchronize(h->c.stream));
  h->profiling = enabled != 0;
  h->c.prof.on = h->profiling;
  h->c.prof.n = 0;
  if (h->profiling)
    for (int i = 0; i < 2 * h->c.prof.cap; ++i)
      if (!h->c.prof.ev[i]) CU(cudaEventCreate(&h->c.prof.ev[i]));
  return 0;
}

int euler_gpu_reset_profile(euler_gpu* h) {
  ENTER(h);
  CU(cudaStreamSynchronize(h->c.stream));
  prof_collect(h->c);
  for (int i = 0; i < KC__COUNT; ++i) { h->c.prof.ms[i] = 0; h->c.prof.count[i] = 0; }
  h->ms_markers = h->ms_grid = h->ms_project = 0;
  return 0;
}

const char* euler_gpu_kernel_class_name(int i) {
  static const char* names[KC__COUNT] = {
      "maxsq", "advect_markers", "refresh_counts", "sources", "extrapolate_bounds",
      "advect_velocity", "build_rhs", "precon_build", "precon_apply", "apply_a", "axpy_norm",
      "update_search", "pressure_update", "misc", "fused_search_apply_a", "fused_axpy_forward",
      "rb_forward", "rb_backward", "color_transport", "true_residual", "fused_tail"};
  return (i >= 0 && i < KC__COUNT) ? names[i] : nullptr;
}

int euler_gpu_trace_read(euler_gpu* h, unsigned long long* out, size_t max_slots, size_t* n_slots) {
  ENTER(h);
  if (!n_slots) return fail(EULER_E_INVALID, "n_slots is NULL");
  Ctx& c = h->c;
  CU(cudaStreamSynchronize(c.stream));
  size_t n = (size_t)c.trace_n;
  if (n > max_slots) n = max_slots;
  if (n && !out) return fail(EULER_E_INVALID, "out is NULL");
  if (n) CU(cudaMemcpy(out, c.trace, n * TRACE_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  *n_slots = n;
  // per-block records of the last traced launch of kernel EULER_TRACE_BLOCKS, behind the slots
  if (c.trace && out && max_slots >= n + (size_t)TRACE_BLK_MAX * TRACE_BLK_WORDS / TRACE_WORDS)
    CU(cudaMemcpy(out + n * TRACE_WORDS, c.trace + (size_t)c.trace_cap * TRACE_WORDS,
                  (size_t)TRACE_BLK_MAX * TRACE_BLK_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  // start over: the next launch records into slot 0 again
  if (c.trace) CU(cudaMemsetAsync(c.trace, 0, (size_t)c.trace_cap * TRACE_WORDS * sizeof(unsigned long long), c.stream));
  c.trace_n = 0;
  return 0;
}

int euler_gpu_synchronize(euler_gpu* h) {
  ENTER(h);
  CU(cudaStreamSynchronize(h->c.stream));
  return check_launch("synchronize");
}

void* euler_gpu_stream(euler_gpu* h) { return h ? (void*)h->c.stream : nullptr; }

int euler_gpu_comm_unique_id(void* unique_id_128) {
  if (!unique_id_128) return fail(EULER_E_INVALID, "unique_id is NULL");
  if (comm_unique_id(unique_id_128)) return fail(EULER_E_COMM, "%s", comm_last_error());
  return 0;
}

int euler_gpu_comm_init(euler_gpu* h, int rank, int n_ranks, const void* unique_id_128) {
  ENTER(h);
  if (!h->slab) return fail(EULER_E_INVALID, "handle was not created with slab_rows > 0");
  if (!unique_id_128 || rank < 0 || rank >= n_ranks) return fail(EULER_E_INVALID, "bad rank/id");
  if (comm_init(&h->cm, rank, n_ranks, unique_id_128)) return fail(EULER_E_COMM, "%s", comm_last_error());
  Ctx& c = h->c;
  void* raw = nullptr;
  CU(cudaMalloc(&raw, sizeof(double) * GATHER_SLOTS * n_ranks)); h->allocs.push_back(raw);
  h->cm.gather = (double*)raw;
  CU(cudaMalloc(&raw, 2 * sizeof(unsigned long long))); h->allocs.push_back(raw);
  h->cm.mig = (unsigned long long*)raw;
  h->cm.send_cap = 32 * (size_t)h->nx;              // markers crossing one boundary per sub-step
  CU(cudaMalloc(&raw, h->cm.send_cap * 8)); h->allocs.push_back(raw); h->cm.send_dn = (float2*)raw;
  CU(cudaMalloc(&raw, h->cm.send_cap * 8)); h->allocs.push_back(raw); h->cm.send_up = (float2*)raw;
  h->device_bytes += 2 * h->cm.send_cap * 8;
  c.distributed = 1;
  h->comm_ready = true;
  { int grc = agree_on_sources(h); if (grc) return grc; }
  // halo rows of the initial classification (create() binned only the owned markers)
  CM(comm_halo(c, h->cm, c.count, 1, SLAB_HALO));
  { int grc = color_halos(h); if (grc) return grc; }
  CU(cudaStreamSynchronize(c.stream));
  return 0;
}

int euler_gpu_comm_p2p_export(euler_gpu* h, void* blob_256) {
  ENTER(h);
  if (!h->slab || !h->comm_ready) return fail(EULER_E_INVALID, "comm_init first");
  if (!blob_256) return fail(EULER_E_INVALID, "blob is NULL");
  if (p2p_export(h->c, h->cm, h->pp, h->z_raw, blob_256)) return fail(EULER_E_COMM, "%s", comm_last_error());
  return 0;
}

int euler_gpu_comm_p2p_import(euler_gpu* h, const void* blobs) {
  ENTER(h);
  if (!h->slab || !h->comm_ready || !h->pp.mine) return fail(EULER_E_INVALID, "comm_p2p_export first");
  if (!blobs) return fail(EULER_E_INVALID, "blobs is NULL");
  CU(cudaStreamSynchronize(h->c.stream));
  if (p2p_import(h->c, h->cm, h->pp, blobs, h->c.mixed ? 4 : 8)) return fail(EULER_E_COMM, "%s", comm_last_error());
  if (h->c.fused == 1) {
    const char* e = getenv("EULER_P2P_SEPARATE");
    h->c.p2p_mode = (e && atoi(e) && !h->c.mixed) ? 1 : 2;
    // split-phase scalar exchange (post in the producing kernel, collect in every block of the
    // consuming one; needs the fused tail kernel).  Measured on the 16384^2 workload, same boxes,
    // back to back (profiles/r02b_*): 56.26 -> 55.94 ms at N = 2, 31.39 -> 31.11 at N = 4,
    // 19.04 -> 18.68 at N = 8, identical results.  EULER_P2P_SPLIT=0 restores the blocking form.
    const char* sp = getenv("EULER_P2P_SPLIT");
    h->split = h->c.p2p_mode == 2 && h->use_tail && !(sp && atoi(sp) == 0);
  }
  return 0;
}

int euler_gpu_slab_partition_weighted(const uint64_t* row_weight, int global_ny, int n_ranks, int rank,
                                      int* row0, int* rows) {
  if (!row_weight || global_ny < 1 || n_ranks < 1 || rank < 0 || rank >= n_ranks || !row0 || !rows)
    return fail(EULER_E_INVALID, "bad slab partition arguments");
  const int min_rows = 2 * SLAB_HALO;
  if (global_ny < n_ranks * min_rows) return fail(EULER_E_INVALID, "grid too short for %d slabs", n_ranks);
  // boundary k sits where the running weight first reaches k/n of the total, then boundaries
  // are pushed apart so that every slab keeps at least min_rows rows
  std::vector<int> cut(n_ranks + 1, 0);
  long double total = 0;
  for (int y = 0; y < global_ny; ++y) total += (long double)row_weight[y];
  cut[n_ranks] = global_ny;
  long double run = 0;
  int k = 1;
  for (int y = 0; y < global_ny && k < n_ranks; ++y) {
    run += (long double)row_weight[y];
    while (k < n_ranks && run * n_ranks >= total * k) cut[k++] = y + 1;
  }
  for (; k < n_ranks; ++k) cut[k] = global_ny;
  for (int i = 1; i < n_ranks; ++i) if (cut[i] < cut[i - 1] + min_rows) cut[i] = cut[i - 1] + min_rows;
  for (int i = n_ranks - 1; i >= 1; --i) if (cut[i] > cut[i + 1] - min_rows) cut[i] = cut[i + 1] - min_rows;
  *row0 = cut[rank];
  *rows = cut[rank + 1] - cut[rank];
  return 0;
}

int euler_gpu_slab_partition(int global_ny, int n_ranks, int rank, int* row0, int* rows) {
  if (global_ny < 1 || n_ranks < 1 || rank < 0 || rank >= n_ranks || !row0 || !rows)
    return fail(EULER_E_INVALID, "bad slab partition arguments");
  // contiguous, balanced to within one row; multiples of nothing in particular: rows are
  // independent units (x is the fast axis, main.c:64)
  const int base = global_ny / n_ranks, extra = global_ny % n_ranks;
  *rows = base + (rank < extra ? 1 : 0);
  *row0 = rank * base + (rank < extra ? rank : extra);
  return 0;
}

}  // extern "C"
