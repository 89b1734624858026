/* euler_b200/host/scenario.c — see scenario.h. */
#include "scenario.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

uint32_t euler_rng_next(uint64_t *state) {
  uint64_t s = *state;
  s ^= s >> 12;
  s ^= s << 25;
  s ^= s >> 27;
  *state = s;
  return (uint32_t)((s * 0x2545F4914F6CDD1Dull) >> 32);
}

float euler_randf(uint64_t *state) {
  return (float)(euler_rng_next(state) / (double)UINT32_MAX);
}

void euler_scenario_free(euler_scenario *s) {
  if (!s) return;
  free(s->solid); free(s->source); free(s->sink); free(s->fluid); free(s->markers);
  memset(s, 0, sizeof *s);
}

int euler_scenario_from_text(euler_scenario *s, const char *text, long length, int nx, int ny) {
  memset(s, 0, sizeof *s);
  const size_t n = (size_t)nx * ny;
  s->nx = nx; s->ny = ny;
  s->solid = calloc(n, 1); s->source = calloc(n, 1); s->sink = calloc(n, 1); s->fluid = calloc(n, 1);
  if (!s->solid || !s->source || !s->sink || !s->fluid) { euler_scenario_free(s); return -1; }
#define AT(p, x, y) (p)[(size_t)(y) * nx + (x)]
  long pos = 0;
  size_t n_fluid = 0;
  for (int y = ny - 2; y > 0 && pos < length; --y) {
    int x = 1;
    for (; x < nx - 1 && pos < length; ++x) {
      const char ch = text[pos++];
      if (ch == '\n') break;
      if (ch == 'X') AT(s->solid, x, y) = 1;
      else if (ch == '0') { AT(s->fluid, x, y) = 1; n_fluid++; }
      else if (ch == '?') { AT(s->fluid, x, y) = 1; AT(s->source, x, y) = 1; n_fluid++; }
      else if (ch == '=') AT(s->sink, x, y) = 1;
    }
    if (x == nx - 1) while (pos < length && text[pos++] != '\n') {}
  }
  for (int y = 0; y < ny; ++y) { AT(s->sink, 0, y) = 1; AT(s->sink, nx - 1, y) = 1; }
  for (int x = 0; x < nx; ++x) { AT(s->sink, x, 0) = 1; AT(s->sink, x, ny - 1) = 1; }

  s->markers = malloc((n_fluid ? n_fluid : 1) * 4 * 2 * sizeof(float));
  if (!s->markers) { euler_scenario_free(s); return -1; }
  uint64_t rng = EULER_RNG_SEED;
  size_t m = 0;
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y) {
      if (!AT(s->fluid, x, y)) continue;
      for (int k = 0; k < 4; ++k) {
        const float jx = euler_randf(&rng) / 2;
        const float jy = euler_randf(&rng) / 2;
        s->markers[2 * m]     = x + (k < 2 ? 0 : 0.5f) + jx;
        s->markers[2 * m + 1] = y + (k % 2 ? 0 : 0.5f) + jy;
        m++;
      }
    }
#undef AT
  s->n_markers = m;
  s->rng_state = rng;
  return 0;
}

int euler_scenario_markers_row_major(euler_scenario *s) {
  const int nx = s->nx, ny = s->ny;
  if (!s->n_markers) return 0;
  float *out = malloc(s->n_markers * 2 * sizeof(float));
  size_t *rank = malloc((size_t)nx * ny * sizeof(size_t));
  if (!out || !rank) { free(out); free(rank); return -1; }
  size_t r = 0;
  for (size_t c = 0; c < (size_t)nx * ny; ++c) { rank[c] = r; r += s->fluid[c] ? 1 : 0; }
  size_t src = 0;                       /* markers were seeded x outermost, y inner, 4 per cell */
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y) {
      const size_t c = (size_t)y * nx + x;
      if (!s->fluid[c]) continue;
      memcpy(out + 8 * rank[c], s->markers + 2 * src, 8 * sizeof(float));
      src += 4;
    }
  free(s->markers);
  free(rank);
  s->markers = out;
  return 0;
}

int euler_scenario_load(euler_scenario *s, const char *path, int nx, int ny) {
  FILE *f = fopen(path, "rb");
  if (!f) return -2;
  if (fseek(f, 0, SEEK_END)) { fclose(f); return -2; }
  long len = ftell(f);
  if (len < 0 || fseek(f, 0, SEEK_SET)) { fclose(f); return -2; }
  char *buf = malloc((size_t)len + 1);
  if (!buf) { fclose(f); return -1; }
  if (len && fread(buf, (size_t)len, 1, f) != 1) { free(buf); fclose(f); return -2; }
  fclose(f);
  int rc = euler_scenario_from_text(s, buf, len, nx, ny);
  free(buf);
  return rc;
}

char *euler_scenario_resample(const char *text, long length, int out_w, int out_h, long *out_len) {
  /* index the input lines */
  long n_lines = 0, max_w = 0;
  for (long i = 0, start = 0; i <= length; ++i)
    if (i == length || text[i] == '\n') {
      if (i == length && i == start) break;
      if (i - start > max_w) max_w = i - start;
      n_lines++; start = i + 1;
    }
  if (n_lines == 0 || max_w == 0) return NULL;
  long *starts = malloc(sizeof(long) * n_lines), *lens = malloc(sizeof(long) * n_lines);
  char *out = malloc((size_t)(out_w + 1) * out_h + 1);
  if (!starts || !lens || !out) { free(starts); free(lens); free(out); return NULL; }
  long k = 0;
  for (long i = 0, start = 0; i <= length && k < n_lines; ++i)
    if (i == length || text[i] == '\n') { starts[k] = start; lens[k] = i - start; k++; start = i + 1; }
  char *w = out;
  for (int j = 0; j < out_h; ++j) {
    const long r = (long)((long long)j * n_lines / out_h);
    for (int i = 0; i < out_w; ++i) {
      const long c = (long)((long long)i * max_w / out_w);
      *w++ = c < lens[r] ? text[starts[r] + c] : ' ';
    }
    *w++ = '\n';
  }
  *w = '\0';
  if (out_len) *out_len = w - out;
  free(starts); free(lens);
  return out;
}

char *euler_scenario_export(int nx, int ny, const uint8_t *solid, const uint8_t *source, const uint8_t *sink,
                            const uint8_t *count, long *out_len) {
  if (nx < 3 || ny < 3) return NULL;
  const int w = nx - 2, h = ny - 2;
  char *out = malloc((size_t)(w + 1) * h + 1);
  if (!out) return NULL;
  char *p = out;
  for (int r = 0; r < h; ++r) {
    const int y = ny - 2 - r;
    for (int i = 0; i < w; ++i) {
      const size_t c = (size_t)y * nx + (size_t)(1 + i);
      *p++ = solid[c] ? 'X' : source[c] ? '?' : sink[c] ? '=' : count[c] ? '0' : ' ';
    }
    *p++ = '\n';
  }
  *p = '\0';
  if (out_len) *out_len = p - out;
  return out;
}

char *euler_scenario_synthetic(const char *name, int nx, int ny, long *out_len) {
  const int w = nx - 2, h = ny - 2;
  int mode;
  if (!strcmp(name, "basic-fill")) mode = 0;
  else if (!strcmp(name, "full")) mode = 1;
  else return NULL;
  char *out = malloc((size_t)(w + 1) * h + 1);
  if (!out) return NULL;
  char *p = out;
  const int fluid_cols = mode == 0 ? (int)(0.4 * w) : w - 2;
  const int fluid_rows = mode == 0 ? (int)(0.5 * h) : h - 2;
  for (int j = 0; j < h; ++j) {           /* j = 0 is the top text row */
    for (int i = 0; i < w; ++i) {
      char ch = ' ';
      if (j == 0 || j == h - 1 || i == 0 || i == w - 1) ch = 'X';
      else if (i <= fluid_cols && j >= h - 1 - fluid_rows) ch = '0';
      *p++ = ch;
    }
    *p++ = '\n';
  }
  *p = '\0';
  if (out_len) *out_len = p - out;
  return out;
}
