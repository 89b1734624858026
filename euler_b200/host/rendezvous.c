/* euler_b200/host/rendezvous.c — see rendezvous.h. */
#define _POSIX_C_SOURCE 200809L
#include "rendezvous.h"

#include <stdio.h>
#include <string.h>
#include <time.h>

static int path_of(char *out, size_t cap, const char *dir, const char *name, int rank, const char *suffix) {
  const int n = snprintf(out, cap, "%s/%s%d%s", dir, name, rank, suffix);
  return n > 0 && (size_t)n < cap ? 0 : -1;
}

int euler_rdv_publish(const char *dir, const char *name, int rank, const void *data, size_t bytes) {
  char tmp[4096], fin[4096];
  if (path_of(tmp, sizeof tmp, dir, name, rank, ".tmp") || path_of(fin, sizeof fin, dir, name, rank, ".bin")) return -1;
  FILE *f = fopen(tmp, "wb");
  if (!f) return -1;
  int bad = bytes && fwrite(data, bytes, 1, f) != 1;
  if (fclose(f)) bad = 1;
  if (bad || rename(tmp, fin)) { remove(tmp); return -1; }
  return 0;
}

int euler_rdv_fetch(const char *dir, const char *name, int rank, void *data, size_t bytes, int timeout_s) {
  char fin[4096];
  if (path_of(fin, sizeof fin, dir, name, rank, ".bin")) return -1;
  struct timespec t0, now;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (;;) {
    FILE *f = fopen(fin, "rb");
    if (f) {                                   /* rename is atomic: the file is complete once visible */
      int bad = fseek(f, 0, SEEK_END) != 0;
      const long size = bad ? -1 : ftell(f);
      bad = bad || size < 0 || (size_t)size != bytes || fseek(f, 0, SEEK_SET) != 0;
      if (!bad && bytes) bad = fread(data, bytes, 1, f) != 1;
      fclose(f);
      return bad ? -1 : 0;
    }
    clock_gettime(CLOCK_MONOTONIC, &now);
    if ((double)(now.tv_sec - t0.tv_sec) + 1e-9 * (double)(now.tv_nsec - t0.tv_nsec) >= (double)timeout_s) return -2;
    const struct timespec nap = {0, 2000000};  /* 2 ms */
    nanosleep(&nap, NULL);
  }
}
