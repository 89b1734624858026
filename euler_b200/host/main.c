/* euler_b200/host/main.c — the host program: CLI, scenario loading, step/render loop.
 *
 * Same user-facing behaviour as the reference's main() (main.c:982-1042): one positional
 * scenario file in the reference's format, keys p (pause) / f (advance one frame while
 * paused) / q (quit), 10 frames per second, 0.1 s of simulated time per frame, ASCII picture
 * of the marker-count plane.  The simulation itself — everything the reference does inside
 * sim_step() — runs on the GPU through the C-ABI of libeuler_gpu.so (include/euler_gpu.h),
 * which this program binds at run time with dlopen; rendering stays on the host and needs
 * only the uint8 count plane per drawn frame.
 *
 * Additions over the reference CLI (it cannot run without a tty, main.c:1005-1007):
 *   --headless          no tty, no pacing; prints one JSON line of statistics at the end
 *   --frames N          stop after N frames (default: run until 'q'; 100 in --headless)
 *   --grid WxH          grid size (default 100x40 as the reference, main.c:22-25); the
 *                       scenario text is resampled (nearest neighbour) when it differs
 *   --synthetic NAME    built-in scenario instead of a file: basic-fill | full
 *   --precon ic0|rb     reference-faithful IC(0) wavefront (default) | red-black IC(0)
 *   --markers ref|fast  reference marker order & dt carry-over (default) | per-marker dt
 *   --exact-dot         reference-order dot products (bit-identical solve)
 *   --pcg-dtype fp64|fp32  storage precision of the PCG vectors: the reference's doubles
 *                       (default) | fp32 planes with fp64 pressure, dot products and residual
 *                       replacement (needs --precon rb; include/euler_gpu.h euler_pcg_dtype)
 *   --device D          CUDA device
 *   --print             in --headless: print the final picture (plain ASCII)
 *   --load / --save F   restore / write a checkpoint of the dynamic state (checkpoint.c)
 *   --export F          write the final state as a scenario file (X 0 ? =; scenario.h)
 *   --ranks N --rank R --rendezvous DIR [--no-p2p]
 *                       one of N processes of a row-slab decomposed run: this
 *                       process drives GPU R (or --device) and owns a slab of rows balanced by
 *                       fluid cells; the NCCL id and the NVLink peer handles are exchanged
 *                       through files in DIR (fresh per run, rendezvous.h).  --headless only;
 *                       implies --precon rb --markers fast.  Rank 0 prints the statistics of
 *                       the whole grid.
 */
#define _POSIX_C_SOURCE 200809L
#include <dlfcn.h>
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "checkpoint.h"
#include "euler_gpu.h"
#include "rendezvous.h"
#include "render.h"
#include "scenario.h"

typedef struct api {
  void *dl;
  int (*default_params)(euler_params *);
  int (*create)(euler_gpu **, int, int, const uint8_t *, const uint8_t *, const uint8_t *,
                const float *, size_t, const euler_params *);
  int (*destroy)(euler_gpu *);
  int (*step_frame)(euler_gpu *, int *);
  int (*read_marker_count)(euler_gpu *, uint8_t *);
  int (*read_window)(euler_gpu *, int, int, int, int, int, void *);
  int (*colorize)(euler_gpu *);
  euler_ckpt_api ck;
  int (*stats)(euler_gpu *, euler_stats *);
  const char *(*last_error)(void);
  /* row slabs (only bound when --ranks is given) */
  int (*slab_partition_weighted)(const uint64_t *, int, int, int, int *, int *);
  int (*comm_unique_id)(void *);
  int (*comm_init)(euler_gpu *, int, int, const void *);
  int (*comm_p2p_export)(euler_gpu *, void *);
  int (*comm_p2p_import)(euler_gpu *, const void *);
} api;

/* What draw_rows() looks at (main.c:917-920): the top `rows` text rows and the left `cols`
 * columns of the interior.  Only that rectangle of the count plane comes back from the GPU. */
static int read_visible(const api *a, euler_gpu *sim, int nx, int ny, const euler_screen *scr, uint8_t *count,
                        float *rgb[3]) {
  int y_low = ny - 1 - scr->rows;
  if (y_low < 1) y_low = 1;
  int w = nx - 2 < scr->cols ? nx - 2 : scr->cols;
  if (w < 0) w = 0;
  int rc = a->read_window(sim, EULER_F_COUNT, 1, y_low, w, ny - 1 - y_low, count);
  if (!rc && rgb[0]) {                                        /* --rainbow: g_r, g_g, g_b (main.c:938) */
    static const int f[3] = {EULER_F_CR, EULER_F_CG, EULER_F_CB};
    for (int k = 0; k < 3 && !rc; ++k) rc = a->read_window(sim, f[k], 1, y_low, w, ny - 1 - y_low, rgb[k]);
  }
  return rc;
}

static void draw_visible(euler_screen *scr, int nx, int ny, const euler_scenario *scn, const uint8_t *count,
                         float *rgb[3]) {
  if (rgb[0]) euler_draw_rainbow(scr, nx, ny, scn->solid, scn->sink, count, rgb[0], rgb[1], rgb[2]);
  else euler_draw(scr, nx, ny, scn->solid, scn->sink, count);
}

static int bind_api(api *a, const char *argv0) {
  const char *env = getenv("EULER_GPU_LIB");
  char path[4096];
  const char *cands[4]; int n = 0;
  if (env) cands[n++] = env;
  /* next to the executable: <root>/bin/euler-gpu -> <root>/euler_b200/lib/libeuler_gpu.so */
  const char *slash = strrchr(argv0, '/');
  if (slash) {
    snprintf(path, sizeof path, "%.*s/../euler_b200/lib/libeuler_gpu.so", (int)(slash - argv0), argv0);
    cands[n++] = path;
  }
  cands[n++] = "euler_b200/lib/libeuler_gpu.so";
  cands[n++] = "libeuler_gpu.so";
  for (int i = 0; i < n && !a->dl; ++i) a->dl = dlopen(cands[i], RTLD_NOW | RTLD_LOCAL);
  if (!a->dl) { fprintf(stderr, "cannot load libeuler_gpu.so: %s\n", dlerror()); return -1; }
#define BIND(field, sym) do { *(void **)(&a->field) = dlsym(a->dl, sym); \
    if (!a->field) { fprintf(stderr, "libeuler_gpu.so lacks %s\n", sym); return -1; } } while (0)
  BIND(default_params, "euler_gpu_default_params");
  BIND(create, "euler_gpu_create");
  BIND(destroy, "euler_gpu_destroy");
  BIND(step_frame, "euler_gpu_step_frame");
  BIND(read_marker_count, "euler_gpu_read_marker_count");
  BIND(read_window, "euler_gpu_read_window");
  BIND(colorize, "euler_gpu_colorize");
  BIND(ck.get, "euler_gpu_get");
  BIND(ck.set, "euler_gpu_set");
  BIND(ck.stats, "euler_gpu_stats");
  BIND(ck.set_rng_state, "euler_gpu_set_rng_state");
  BIND(ck.set_source_exhausted, "euler_gpu_set_source_exhausted");
  BIND(ck.set_frame_count, "euler_gpu_set_frame_count");
  BIND(stats, "euler_gpu_stats");
  BIND(last_error, "euler_gpu_last_error");
  BIND(slab_partition_weighted, "euler_gpu_slab_partition_weighted");
  BIND(comm_unique_id, "euler_gpu_comm_unique_id");
  BIND(comm_init, "euler_gpu_comm_init");
  BIND(comm_p2p_export, "euler_gpu_comm_p2p_export");
  BIND(comm_p2p_import, "euler_gpu_comm_p2p_import");
#undef BIND
  return 0;
}

static uint64_t fnv1a(const uint8_t *p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

static void usage(const char *argv0) {
  fprintf(stderr,
          "usage: %s [--rainbow] [--headless] [--frames N] [--grid WxH] [--synthetic NAME]\n"
          "       [--precon ic0|rb] [--markers ref|fast] [--exact-dot] [--pcg-dtype fp64|fp32]\n"
          "       [--device D] [--print]\n"
          "       [--load CHECKPOINT] [--save CHECKPOINT] [--export SCENARIO]\n"
          "       [--ranks N --rank R --rendezvous DIR [--no-p2p]] <scenario>\n",
          argv0);
}

/* Argument syntax, checked BEFORE the CUDA library is bound (usage and "Unrecognized input" errors
 * must not need a loadable libeuler_gpu.so): unknown options, options without their value and
 * more than one scenario file are errors, like the reference's parse_args (main.c:982-1002), which
 * rejects every token it does not know. */
static int check_args(int argc, char **argv) {
  static const char *const flags[] = {"--rainbow", "--headless", "--print", "--exact-dot", "--no-p2p", NULL};
  static const char *const valued[] = {"--frames", "--device", "--ranks", "--rank", "--rendezvous", "--synthetic",
                                       "--load", "--save", "--export", "--grid", "--precon", "--pcg-dtype",
                                       "--markers", NULL};
  int files = 0, synthetic = 0;
  for (int i = 1; i < argc; ++i) {
    const char *s = argv[i];
    int known = 0;
    for (int k = 0; flags[k] && !known; ++k) known = !strcmp(s, flags[k]);
    for (int k = 0; valued[k] && !known; ++k)
      if (!strcmp(s, valued[k])) {
        if (i + 1 >= argc) { usage(argv[0]); return 1; }
        if (!strcmp(s, "--synthetic")) synthetic = 1;
        ++i; known = 1;
      }
    if (known) continue;
    if (s[0] == '-' && s[1] == '-') { fprintf(stderr, "Unrecognized input: %s\n", s); return 1; }   /* main.c:995 */
    if (++files > 1) { fprintf(stderr, "Unrecognized input: %s\n", s); return 1; }
  }
  if (!files && !synthetic) { usage(argv[0]); return 1; }                                            /* main.c:986-989 */
  return 0;
}

int main(int argc, char **argv) {
  if (check_args(argc, argv)) return 1;
  const char *file = NULL, *synthetic = NULL, *load_path = NULL, *save_path = NULL, *export_path = NULL;
  int headless = 0, frames = -1, nx = 100, ny = 40, do_print = 0;
  int ranks = 1, rank = 0, no_p2p = 0, device_given = 0;
  const char *rdv = NULL;
  api a; memset(&a, 0, sizeof a);
  if (bind_api(&a, argv[0])) return 1;
  euler_params prm;
  a.default_params(&prm);

  for (int i = 1; i < argc; ++i) {
    const char *s = argv[i];
    if (!strcmp(s, "--rainbow")) prm.rainbow = 1;                /* main.c:991-992 */
    else if (!strcmp(s, "--headless")) headless = 1;
    else if (!strcmp(s, "--print")) do_print = 1;
    else if (!strcmp(s, "--exact-dot")) prm.dot_mode = EULER_DOT_REFERENCE_ORDER;
    else if (!strcmp(s, "--frames") && i + 1 < argc) frames = atoi(argv[++i]);
    else if (!strcmp(s, "--device") && i + 1 < argc) { prm.device = atoi(argv[++i]); device_given = 1; }
    else if (!strcmp(s, "--ranks") && i + 1 < argc) ranks = atoi(argv[++i]);
    else if (!strcmp(s, "--rank") && i + 1 < argc) rank = atoi(argv[++i]);
    else if (!strcmp(s, "--rendezvous") && i + 1 < argc) rdv = argv[++i];
    else if (!strcmp(s, "--no-p2p")) no_p2p = 1;
    else if (!strcmp(s, "--synthetic") && i + 1 < argc) synthetic = argv[++i];
    else if (!strcmp(s, "--load") && i + 1 < argc) load_path = argv[++i];
    else if (!strcmp(s, "--save") && i + 1 < argc) save_path = argv[++i];
    else if (!strcmp(s, "--export") && i + 1 < argc) export_path = argv[++i];
    else if (!strcmp(s, "--grid") && i + 1 < argc) {
      if (sscanf(argv[++i], "%dx%d", &nx, &ny) != 2 || nx < 4 || ny < 4) { usage(argv[0]); return 1; }
    } else if (!strcmp(s, "--precon") && i + 1 < argc) {
      const char *v = argv[++i];
      if (!strcmp(v, "ic0")) prm.precon = EULER_PRECON_IC0_WAVEFRONT;
      else if (!strcmp(v, "rb")) prm.precon = EULER_PRECON_REDBLACK;
      else { usage(argv[0]); return 1; }
    } else if (!strcmp(s, "--pcg-dtype") && i + 1 < argc) {
      const char *v = argv[++i];
      if (!strcmp(v, "fp64")) prm.pcg_dtype = EULER_PCG_FP64;
      else if (!strcmp(v, "fp32")) prm.pcg_dtype = EULER_PCG_FP32;   /* needs --precon rb */
      else { usage(argv[0]); return 1; }
    } else if (!strcmp(s, "--markers") && i + 1 < argc) {
      const char *v = argv[++i];
      if (!strcmp(v, "ref")) prm.marker_mode = EULER_MARKERS_REFERENCE;
      else if (!strcmp(v, "fast")) prm.marker_mode = EULER_MARKERS_FAST;
      else { usage(argv[0]); return 1; }
    } else if (s[0] == '-' && s[1] == '-') {
      fprintf(stderr, "Unrecognized input: %s\n", s);        /* as main.c:995 */
      return 1;
    } else file = s;
  }
  if (!file && !synthetic) { usage(argv[0]); return 1; }    /* as main.c:986-989 */
  if (headless && frames < 0) frames = 100;
  if (ranks > 1) {
    if (!headless || !rdv || rank < 0 || rank >= ranks || load_path || save_path || prm.rainbow) {
      /* (the library transports --rainbow colours on slabs; this program's rank-0 report only assembles the count plane) */
      fprintf(stderr, "--ranks needs --headless, --rank 0..N-1 and --rendezvous DIR (no --load/--save/--rainbow)\n");
      return 1;
    }
    prm.precon = EULER_PRECON_REDBLACK;                       /* the only modes that decompose */
    prm.marker_mode = EULER_MARKERS_FAST;
    if (!device_given) prm.device = rank;
  }

  /* scenario text -> masks + seeded markers (host, reference format and RNG stream) */
  euler_scenario scn;
  int rc;
  if (synthetic) {
    long len = 0;
    char *text = euler_scenario_synthetic(synthetic, nx, ny, &len);
    if (!text) { fprintf(stderr, "unknown synthetic scenario %s\n", synthetic); return 1; }
    rc = euler_scenario_from_text(&scn, text, len, nx, ny);
    free(text);
  } else if (nx == 100 && ny == 40) {
    rc = euler_scenario_load(&scn, file, nx, ny);
  } else {
    FILE *f = fopen(file, "rb");
    rc = -2;
    if (f) {
      long len = -1;
      if (fseek(f, 0, SEEK_END) == 0) len = ftell(f);
      char *raw = (len >= 0 && fseek(f, 0, SEEK_SET) == 0) ? malloc((size_t)len + 1) : NULL;
      if (raw && (len == 0 || fread(raw, (size_t)len, 1, f) == 1)) {
        long rlen = 0;
        char *text = euler_scenario_resample(raw, len, nx - 2, ny - 2, &rlen);
        rc = text ? euler_scenario_from_text(&scn, text, rlen, nx, ny) : -1;
        free(text);
      }
      free(raw); fclose(f);
    }
  }
  if (rc) { fprintf(stderr, "Could not load %s!\n", file ? file : synthetic); return 1; }  /* main.c:213 */

  if (prm.precon == EULER_PRECON_IC0_WAVEFRONT && (long)nx * ny > 512l * 512l)
    fprintf(stderr, "note: --precon ic0 (the default: the reference's natural-order IC(0), solved by a wavefront) is the\n"
                    "      parity mode and latency-bound; on a %dx%d grid --precon rb (red-black IC(0)) is ~10-100x faster\n", nx, ny);
  prm.rng_state = scn.rng_state;
  if (ranks > 1) {
    /* slabs balanced by work: the PCG and the grid stages stream the tiles that hold or border
     * fluid, markers live in fluid cells; a dry row costs one pass over its count bytes (the
     * weights bench.py uses) */
    uint64_t *weight = malloc((size_t)ny * sizeof *weight);
    if (!weight) return 1;
    for (int y = 0; y < ny; ++y) {
      uint64_t wet = 0;
      for (int x = 0; x < nx; ++x) wet += scn.fluid[(size_t)y * nx + x] ? 1 : 0;
      weight[y] = wet * 4096 + (uint64_t)(nx / 256 > 0 ? nx / 256 : 1);
    }
    const int prc = a.slab_partition_weighted(weight, ny, ranks, rank, &prm.slab_row0, &prm.slab_rows);
    free(weight);
    if (prc) { fprintf(stderr, "slab partition: %s\n", a.last_error()); return 1; }
  }
  euler_gpu *sim = NULL;
  if (a.create(&sim, nx, ny, scn.solid, scn.source, scn.sink, scn.markers, scn.n_markers, &prm)) {
    fprintf(stderr, "euler_gpu_create: %s\n", a.last_error());
    return 1;
  }
  uint8_t *count = calloc((size_t)nx * ny, 1);     /* a slab handle fills in only the rows it owns */
  uint64_t run_key = 0;                             /* of this run's rendezvous files */
  if (!count) return 1;
  if (ranks > 1) {
    /* communicator id from rank 0, then (NVLink path) everybody's peer handles, through DIR */
    unsigned char uid[128];
    if (rank == 0 && a.comm_unique_id(uid)) { fprintf(stderr, "cannot make the communicator id: %s\n", a.last_error()); return 1; }
    /* token-echo handshake (rendezvous.h): DIR may hold files of earlier runs */
    const int hrc = euler_rdv_handshake(rdv, rank, ranks, uid, sizeof uid, &run_key, 120);
    if (hrc) {
      fprintf(stderr, hrc == -2 ? "rendezvous in %s timed out: not all %d ranks showed up\n" : "rendezvous in %s failed (I/O)\n", rdv, ranks);
      return 1;
    }
    if (a.comm_init(sim, rank, ranks, uid)) { fprintf(stderr, "comm_init: %s\n", a.last_error()); return 1; }
    if (!no_p2p) {
      unsigned char *blobs = malloc((size_t)ranks * 256);
      if (!blobs) return 1;
      int bad = a.comm_p2p_export(sim, blobs + (size_t)rank * 256) != 0;
      if (!bad) bad = euler_rdv_publish_keyed(rdv, "blob", rank, run_key, blobs + (size_t)rank * 256, 256) != 0;
      for (int r = 0; r < ranks && !bad; ++r) bad = euler_rdv_fetch_keyed(rdv, "blob", r, run_key, blobs + (size_t)r * 256, 256, 120) != 0;
      if (!bad) bad = a.comm_p2p_import(sim, blobs) != 0;
      free(blobs);
      if (bad) { fprintf(stderr, "peer-to-peer set-up failed: %s\n", a.last_error()); return 1; }
    }
  }
  if (load_path) {            /* state of an earlier run; the scenario still supplies the static masks */
    int lrc = euler_checkpoint_load(&a.ck, sim, nx, ny, prm.rainbow, load_path);
    if (lrc) { fprintf(stderr, "cannot load checkpoint %s (%d): %s\n", load_path, lrc, lrc > 0 ? a.last_error() : "bad file"); return 1; }
  }

  long long substeps_total = 0;
  euler_time t0 = euler_now();
  if (headless) {
    for (int f = 0; f < frames; ++f) {
      int sub = 0;
      if (a.step_frame(sim, &sub)) { fprintf(stderr, "step: %s\n", a.last_error()); return 1; }
      substeps_total += sub;
    }
    if (a.read_marker_count(sim, count)) { fprintf(stderr, "read: %s\n", a.last_error()); return 1; }
  } else {
    euler_screen scr; memset(&scr, 0, sizeof scr);
    if (euler_tty_window_size(&scr.rows, &scr.cols) == -1) {
      fprintf(stderr, "stdout is not a terminal: use --headless\n");
      return 1;
    }
    euler_tty_raw_mode();
    euler_tty_clear();
    memset(count, 0, (size_t)nx * ny);
    float *rgb[3] = {NULL, NULL, NULL};
    if (prm.rainbow)
      for (int k = 0; k < 3; ++k)
        if (!(rgb[k] = calloc((size_t)nx * ny, sizeof(float)))) return 1;   /* only the window is ever touched */
    read_visible(&a, sim, nx, ny, &scr, count, rgb);
    draw_visible(&scr, nx, ny, &scn, count, rgb);
    int pause = 0, pending = 0, done = 0, f = 0;
    euler_time start = euler_now();
    while (!done && (frames < 0 || f < frames)) {
      const char key = euler_tty_read_key();                 /* main.c:961-980 */
      if (key == 'p') pause = !pause;
      else if (key == 'f') pending++;
      else if (key == 'r') { if (prm.rainbow) a.colorize(sim); }   /* main.c:971-974 */
      else if (key == 'q') { done = 1; break; }
      if (!pause || pending) {                               /* main.c:844-846, 896-898 */
        int sub = 0;
        if (a.step_frame(sim, &sub)) { euler_tty_restore(); fprintf(stderr, "step: %s\n", a.last_error()); return 1; }
        substeps_total += sub; f++;
        if (pending) pending--;
      }
      start = euler_wait_until(start, 100000000ll);          /* 10 fps, main.c:1036 */
      euler_tty_window_size(&scr.rows, &scr.cols);
      read_visible(&a, sim, nx, ny, &scr, count, rgb);
      draw_visible(&scr, nx, ny, &scn, count, rgb);
    }
    for (int k = 0; k < 3; ++k) free(rgb[k]);
    euler_tty_clear();
    euler_tty_restore();
    euler_screen_free(&scr);
    frames = f;
  }
  const double secs = (double)(euler_now().ns - t0.ns) * 1e-9;

  euler_stats st;
  a.stats(sim, &st);
  if (ranks > 1) {
    /* every rank publishes the rows it owns; rank 0 assembles the whole count plane and the
     * global marker count, the others are done */
    const size_t row0 = (size_t)prm.slab_row0 * nx, nrow = (size_t)prm.slab_rows * nx;
    struct { int32_t row0, rows; uint64_t markers; } info = {prm.slab_row0, prm.slab_rows, st.n_markers};
    if (euler_rdv_publish_keyed(rdv, "info", rank, run_key, &info, sizeof info) ||
        euler_rdv_publish_keyed(rdv, "count", rank, run_key, count + row0, nrow)) {
      fprintf(stderr, "cannot publish the results of rank %d\n", rank);
      return 1;
    }
    if (rank != 0) { free(count); a.destroy(sim); euler_scenario_free(&scn); return 0; }
    for (int r = 1; r < ranks; ++r) {
      if (euler_rdv_fetch_keyed(rdv, "info", r, run_key, &info, sizeof info, 600) ||
          info.row0 < 0 || info.rows < 0 || info.row0 + info.rows > ny ||
          euler_rdv_fetch_keyed(rdv, "count", r, run_key, count + (size_t)info.row0 * nx, (size_t)info.rows * nx, 600)) {
        fprintf(stderr, "rank %d never delivered its rows\n", r);
        return 1;
      }
      st.n_markers += info.markers;
    }
  }
  if (do_print) {
    size_t cap = (size_t)(nx + 1) * ny + 1;
    char *pic = malloc(cap);
    if (pic) { euler_draw_plain(pic, cap, nx, ny, 200, 60, scn.solid, scn.sink, count); fputs(pic, stdout); free(pic); }
  }
  char rainbow_json[160] = "";
  if (headless && prm.rainbow) {
    /* FNV-1a of g_r, g_g, g_b masked to the fluid cells (the ones draw_rows reads, main.c:938):
     * the quantity tests/golden/rainbow_answers.json holds from the reference */
    static const int fld[3] = {EULER_F_CR, EULER_F_CG, EULER_F_CB};
    uint64_t hsh[3] = {0, 0, 0};
    float *plane = malloc((size_t)nx * ny * sizeof(float));
    if (!plane) return 1;
    for (int k = 0; k < 3; ++k) {
      if (a.read_window(sim, fld[k], 0, 0, nx, ny, plane)) { fprintf(stderr, "read: %s\n", a.last_error()); return 1; }
      for (size_t i = 0; i < (size_t)nx * ny; ++i) if (!count[i]) plane[i] = 0.f;
      hsh[k] = fnv1a((const uint8_t *)plane, (size_t)nx * ny * sizeof(float));
    }
    free(plane);
    snprintf(rainbow_json, sizeof rainbow_json, ", \"fnv_r\": \"%016" PRIx64 "\", \"fnv_g\": \"%016" PRIx64
             "\", \"fnv_b\": \"%016" PRIx64 "\"", hsh[0], hsh[1], hsh[2]);
  }
  if (headless) {
    printf("{\"grid\": [%d, %d], \"frames\": %d, \"substeps\": %lld, \"seconds\": %.6f, "
           "\"cell_updates_per_s\": %.6e, \"pcg_iterations\": %" PRIu64 ", \"pcg_iters_per_s\": %.6e, "
           "\"solves\": %" PRIu64 ", \"solves_skipped\": %" PRIu64 ", \"markers\": %" PRIu64 ", "
           "\"fnv_count\": \"%016" PRIx64 "\", \"rng_state\": \"%016" PRIx64 "\", \"kernel_launches\": %" PRIu64 "%s}\n",
           nx, ny, frames, substeps_total, secs, (double)nx * ny * (double)substeps_total / secs,
           st.pcg_iterations, (double)st.pcg_iterations / secs, st.solves, st.solves_skipped,
           st.n_markers, fnv1a(count, (size_t)nx * ny), st.rng_state, st.kernel_launches, rainbow_json);
  }
  if (save_path) {
    int src = euler_checkpoint_save(&a.ck, sim, nx, ny, prm.rainbow, save_path);
    if (src) { fprintf(stderr, "cannot save checkpoint %s (%d): %s\n", save_path, src, src > 0 ? a.last_error() : "I/O"); return 1; }
  }
  if (export_path) {            /* the state the last frame left, in the scenario-file format */
    long len = 0;
    char *text = NULL;
    FILE *f = NULL;
    int bad = a.read_marker_count(sim, count) != 0;
    if (!bad) bad = !(text = euler_scenario_export(nx, ny, scn.solid, scn.source, scn.sink, count, &len));
    if (!bad) bad = !(f = fopen(export_path, "wb"));
    if (!bad) bad = len && fwrite(text, (size_t)len, 1, f) != 1;
    if (f && fclose(f)) bad = 1;
    free(text);
    if (bad) { fprintf(stderr, "cannot export scenario %s\n", export_path); return 1; }
  }
  free(count);
  a.destroy(sim);
  euler_scenario_free(&scn);
  return 0;
}
