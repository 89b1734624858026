/* euler_b200/host/rendezvous.c — see rendezvous.h. */
#define _POSIX_C_SOURCE 200809L
#include "rendezvous.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#define RDV_MAX_RANKS 64
#define RDV_MAX_UID 256

static int path_of(char *out, size_t cap, const char *dir, const char *name, int rank, const char *suffix) {
  const int n = snprintf(out, cap, "%s/%s%d%s", dir, name, rank, suffix);
  return n > 0 && (size_t)n < cap ? 0 : -1;
}

static double seconds_since(const struct timespec *t0) {
  struct timespec now;
  clock_gettime(CLOCK_MONOTONIC, &now);
  return (double)(now.tv_sec - t0->tv_sec) + 1e-9 * (double)(now.tv_nsec - t0->tv_nsec);
}

static void nap(void) {
  const struct timespec t = {0, 2000000};      /* 2 ms */
  nanosleep(&t, NULL);
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

/* one attempt: 0 read, 1 not there, -1 there but unreadable / wrong size */
static int try_read(const char *path, void *data, size_t bytes) {
  FILE *f = fopen(path, "rb");
  if (!f) return 1;                            /* rename is atomic: a file is complete once visible */
  int bad = fseek(f, 0, SEEK_END) != 0;
  const long size = bad ? -1 : ftell(f);
  bad = bad || size < 0 || (size_t)size != bytes || fseek(f, 0, SEEK_SET) != 0;
  if (!bad && bytes) bad = fread(data, bytes, 1, f) != 1;
  fclose(f);
  return bad ? -1 : 0;
}

int euler_rdv_fetch(const char *dir, const char *name, int rank, void *data, size_t bytes, int timeout_s) {
  char fin[4096];
  if (path_of(fin, sizeof fin, dir, name, rank, ".bin")) return -1;
  struct timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (;;) {
    const int rc = try_read(fin, data, bytes);
    if (rc <= 0) return rc;
    if (seconds_since(&t0) >= (double)timeout_s) return -2;
    nap();
  }
}

static uint64_t fnv64(const void *p, size_t n, uint64_t h) {
  const unsigned char *b = p;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
  return h;
}

static uint64_t fresh_token(int rank) {
  uint64_t t = 0;
  FILE *f = fopen("/dev/urandom", "rb");
  if (f) { if (fread(&t, sizeof t, 1, f) != 1) t = 0; fclose(f); }
  struct timespec now;
  clock_gettime(CLOCK_REALTIME, &now);
  const uint64_t mix[4] = {t, (uint64_t)now.tv_sec, (uint64_t)now.tv_nsec, ((uint64_t)getpid() << 16) ^ (uint64_t)rank};
  t = fnv64(mix, sizeof mix, 0xcbf29ce484222325ull);
  return t ? t : 1;
}

int euler_rdv_publish_keyed(const char *dir, const char *name, int rank, uint64_t key, const void *data, size_t bytes) {
  unsigned char *buf = malloc(bytes + 8);
  if (!buf) return -1;
  memcpy(buf, &key, 8);
  if (bytes) memcpy(buf + 8, data, bytes);
  const int rc = euler_rdv_publish(dir, name, rank, buf, bytes + 8);
  free(buf);
  return rc;
}

int euler_rdv_fetch_keyed(const char *dir, const char *name, int rank, uint64_t key, void *data, size_t bytes,
                          int timeout_s) {
  char fin[4096];
  if (path_of(fin, sizeof fin, dir, name, rank, ".bin")) return -1;
  unsigned char *buf = malloc(bytes + 8);
  if (!buf) return -1;
  struct timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int rc;
  for (;;) {
    /* a file of another size or with another key is an earlier run's: keep waiting for ours */
    if (try_read(fin, buf, bytes + 8) == 0 && !memcmp(buf, &key, 8)) { if (bytes) memcpy(data, buf + 8, bytes); rc = 0; break; }
    if (seconds_since(&t0) >= (double)timeout_s) { rc = -2; break; }
    nap();
  }
  free(buf);
  return rc;
}

/* answer of rank 0: the tokens it saw (its own at [0]) followed by the communicator id */
typedef struct { uint64_t token[RDV_MAX_RANKS]; unsigned char uid[RDV_MAX_UID]; } rdv_answer;

static uint64_t ack_of(uint64_t token, const void *uid, size_t uid_bytes) {
  return fnv64(uid, uid_bytes, fnv64(&token, 8, 0xcbf29ce484222325ull));
}

int euler_rdv_handshake(const char *dir, int rank, int ranks, void *uid, size_t uid_bytes, uint64_t *key,
                        int timeout_s) {
  if (!dir || !uid || !key || ranks < 1 || ranks > RDV_MAX_RANKS || rank < 0 || rank >= ranks ||
      uid_bytes == 0 || uid_bytes > RDV_MAX_UID) return -1;
  char path[4096];
  struct timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  const uint64_t mine = fresh_token(rank);
  rdv_answer ans;
  memset(&ans, 0, sizeof ans);
  if (rank != 0) {
    if (euler_rdv_publish(dir, "hello", rank, &mine, sizeof mine)) return -1;
    if (path_of(path, sizeof path, dir, "answer", 0, ".bin")) return -1;
    for (;;) {
      if (try_read(path, &ans, sizeof ans) == 0 && ans.token[rank] == mine) break;   /* my token of THIS run echoed */
      if (seconds_since(&t0) >= (double)timeout_s) return -2;
      nap();
    }
    memcpy(uid, ans.uid, uid_bytes);
    const uint64_t ack = ack_of(mine, uid, uid_bytes);
    if (euler_rdv_publish(dir, "ack", rank, &ack, sizeof ack)) return -1;
  } else {
    ans.token[0] = mine;
    memcpy(ans.uid, uid, uid_bytes);
    int published = 0;
    for (;;) {
      int changed = !published, all = 1;
      for (int r = 1; r < ranks; ++r) {
        uint64_t t = 0;
        if (path_of(path, sizeof path, dir, "hello", r, ".bin")) return -1;
        if (try_read(path, &t, sizeof t) != 0) { all = 0; continue; }
        if (t != ans.token[r]) { ans.token[r] = t; changed = 1; }
      }
      if (all && changed) {
        if (euler_rdv_publish(dir, "answer", 0, &ans, sizeof ans)) return -1;
        published = 1;
      }
      int acked = all && published;
      for (int r = 1; r < ranks && acked; ++r) {
        uint64_t a = 0;
        if (path_of(path, sizeof path, dir, "ack", r, ".bin")) return -1;
        acked = try_read(path, &a, sizeof a) == 0 && a == ack_of(ans.token[r], uid, uid_bytes);
      }
      if (acked) break;
      if (seconds_since(&t0) >= (double)timeout_s) return -2;
      nap();
    }
  }
  *key = fnv64(uid, uid_bytes, 0xcbf29ce484222325ull);
  return 0;
}
