/* euler_b200/host/checkpoint.c — see checkpoint.h. */
#include "checkpoint.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const char MAGIC[8] = {'E', 'U', 'L', 'E', 'R', 'C', 'K', '1'};

typedef struct header {
  char magic[8];
  int32_t nx, ny;
  uint32_t flags;
  uint32_t mask_hash_lo;         /* FNV-1a 64 of the solid, source and sink planes the state belongs to */
  uint64_t n_markers, rng_state, frames;
  int32_t source_exhausted;
  uint32_t mask_hash_hi;         /* (0/0 in files written before the hash existed: not checked) */
} header;

typedef struct plane_desc { int field; size_t elem; } plane_desc;
static const plane_desc BASE[] = {{EULER_F_U, 4}, {EULER_F_V, 4}, {EULER_F_COUNT, 1}, {EULER_F_PREV_COUNT, 1},
                                  {EULER_F_PRECON, 8}};
static const plane_desc COLOR[] = {{EULER_F_CR, 4}, {EULER_F_CG, 4}, {EULER_F_CB, 4}};
enum { FLAG_COLOR = 1u, FLAG_NO_PRECON = 2u };

/* A handle created with pcg_dtype = FP32 has no fp64 g_precon plane (its red-black factor is
 * rebuilt from scratch every solve, nothing persists): get/set(EULER_F_PRECON) answer
 * EULER_E_INVALID there.  The file keeps its fixed layout — zeros are written in that case —
 * and bit 1 of the flags says so. */
static int has_precon_plane(const euler_ckpt_api *a, euler_gpu *sim, void *buf, size_t cells) {
  return a->get(sim, EULER_F_PRECON, buf, cells * 8) == 0;
}

/* The dynamic state only makes sense on the static masks it was computed with (markers inside
 * another scenario's solids, sources elsewhere): the header carries their hash. */
static int mask_hash(const euler_ckpt_api *a, euler_gpu *sim, void *buf, size_t cells, uint64_t *out) {
  static const int planes[3] = {EULER_F_SOLID, EULER_F_SOURCE, EULER_F_SINK};
  uint64_t h = 0xcbf29ce484222325ull;
  for (int i = 0; i < 3; ++i) {
    const int rc = a->get(sim, planes[i], buf, cells);
    if (rc) return rc;
    const unsigned char *b = buf;
    for (size_t k = 0; k < cells; ++k) { h ^= b[k]; h *= 0x100000001b3ull; }
  }
  *out = h ? h : 1;
  return 0;
}

int euler_checkpoint_save(const euler_ckpt_api *a, euler_gpu *sim, int nx, int ny, int rainbow, const char *path) {
  euler_stats st;
  int rc = a->stats(sim, &st);
  if (rc) return -rc;
  header h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, MAGIC, 8);
  h.nx = nx; h.ny = ny; h.flags = rainbow ? FLAG_COLOR : 0u;
  h.n_markers = st.n_markers; h.rng_state = st.rng_state; h.frames = st.frames;
  h.source_exhausted = st.source_exhausted;
  const size_t cells = (size_t)nx * ny;
  size_t big = cells * 8 > h.n_markers * 8 ? cells * 8 : (size_t)h.n_markers * 8;
  void *buf = malloc(big ? big : 1);
  if (!buf) return -1;
  const int precon = has_precon_plane(a, sim, buf, cells);
  if (!precon) h.flags |= FLAG_NO_PRECON;
  uint64_t mh = 0;
  if ((rc = mask_hash(a, sim, buf, cells, &mh))) { free(buf); return -rc; }
  h.mask_hash_lo = (uint32_t)mh; h.mask_hash_hi = (uint32_t)(mh >> 32);
  FILE *f = fopen(path, "wb");
  if (!f) { free(buf); return -1; }
  int err = fwrite(&h, sizeof h, 1, f) != 1;
  for (size_t i = 0; !err && i < sizeof BASE / sizeof BASE[0]; ++i) {
    if (BASE[i].field == EULER_F_PRECON && !precon) memset(buf, 0, cells * 8);
    else if ((rc = a->get(sim, BASE[i].field, buf, cells * BASE[i].elem))) break;
    err = fwrite(buf, BASE[i].elem, cells, f) != cells;
  }
  for (size_t i = 0; !err && !rc && rainbow && i < 3; ++i) {
    if ((rc = a->get(sim, COLOR[i].field, buf, cells * 4))) break;
    err = fwrite(buf, 4, cells, f) != cells;
  }
  if (!err && !rc && h.n_markers) {
    if (!(rc = a->get(sim, EULER_F_MARKERS, buf, (size_t)h.n_markers * 8)))
      err = fwrite(buf, 8, (size_t)h.n_markers, f) != (size_t)h.n_markers;
  }
  if (fclose(f)) err = 1;
  free(buf);
  return rc ? -rc : (err ? -1 : 0);
}

int euler_checkpoint_load(const euler_ckpt_api *a, euler_gpu *sim, int nx, int ny, int rainbow, const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  header h;
  if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, MAGIC, 8)) { fclose(f); return -1; }
  if (h.nx != nx || h.ny != ny || ((h.flags & FLAG_COLOR) != 0) != (rainbow != 0) ||
      h.n_markers > 4ull * (uint64_t)nx * (uint64_t)ny) { fclose(f); return -2; }
  const size_t cells = (size_t)nx * ny;
  size_t big = cells * 8 > h.n_markers * 8 ? cells * 8 : (size_t)h.n_markers * 8;
  void *buf = malloc(big ? big : 1);
  if (!buf) { fclose(f); return -1; }
  int rc = 0, err = 0;
  const uint64_t want = ((uint64_t)h.mask_hash_hi << 32) | h.mask_hash_lo;
  if (want) {                                  /* a state of another scenario of the same size: refuse */
    uint64_t mh = 0;
    if ((rc = mask_hash(a, sim, buf, cells, &mh))) { free(buf); fclose(f); return -rc; }
    if (mh != want) { free(buf); fclose(f); return -2; }
  }
  /* the plane is restored when both the file and the handle have it; a handle that has it while
   * the file does not starts from the zero-initialised g_precon of a fresh run (main.c:577): the
   * IC(0) iterates depend on what the plane holds at non-fluid cells */
  const int handle_precon = has_precon_plane(a, sim, buf, cells);
  const int precon = !(h.flags & FLAG_NO_PRECON) && handle_precon;
  for (size_t i = 0; !err && !rc && i < sizeof BASE / sizeof BASE[0]; ++i) {
    err = fread(buf, BASE[i].elem, cells, f) != cells;
    if (BASE[i].field == EULER_F_PRECON && !precon) {
      if (!err && handle_precon) { memset(buf, 0, cells * 8); rc = a->set(sim, EULER_F_PRECON, buf, cells * 8); }
      continue;
    }
    if (!err) rc = a->set(sim, BASE[i].field, buf, cells * BASE[i].elem);
  }
  for (size_t i = 0; !err && !rc && rainbow && i < 3; ++i) {
    err = fread(buf, 4, cells, f) != cells;
    if (!err) rc = a->set(sim, COLOR[i].field, buf, cells * 4);
  }
  if (!err && !rc) {
    if (h.n_markers) err = fread(buf, 8, (size_t)h.n_markers, f) != (size_t)h.n_markers;
    if (!err) rc = a->set(sim, EULER_F_MARKERS, buf, (size_t)h.n_markers * 8);
  }
  if (!err && !rc) rc = a->set_rng_state(sim, h.rng_state);
  if (!err && !rc) rc = a->set_source_exhausted(sim, h.source_exhausted);
  if (!err && !rc) rc = a->set_frame_count(sim, h.frames);
  fclose(f);
  free(buf);
  return rc ? -rc : (err ? -1 : 0);
}
