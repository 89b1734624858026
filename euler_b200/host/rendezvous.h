/* euler_b200/host/rendezvous.h — file rendezvous between the N processes of one slab-decomposed
 * run (one process per GPU).  The C-ABI leaves the side channel to the caller
 * (euler_gpu_comm_init takes the 128-byte communicator id "distributed over any side channel");
 * bench.py uses torch.distributed for it, the host C program uses a directory the ranks share:
 * rank r publishes <dir>/<name><r>.bin atomically (write to a temporary name, then rename) and
 * the others poll for it.
 *
 * The directory does NOT have to be fresh.  euler_rdv_handshake proves freshness without any
 * shared secret: every rank publishes a random token, rank 0 answers with the communicator id
 * next to the tokens it saw, a rank accepts the id only when its OWN token of THIS run is echoed
 * (a stale answer of an earlier run carries other tokens and is ignored; rank 0 re-reads the
 * tokens and re-publishes until every rank has acknowledged the id it published).  The handshake
 * yields a 64-bit run key (a hash of the id, which is unique per run) that is stored in front of
 * every later file; euler_rdv_fetch_keyed ignores files that carry another key, so results of
 * an earlier run left in the directory can never be mistaken for this run's. */
#ifndef EULER_RENDEZVOUS_H
#define EULER_RENDEZVOUS_H
#include <stddef.h>
#include <stdint.h>

/* 0, or -1 on an I/O error */
int euler_rdv_publish(const char *dir, const char *name, int rank, const void *data, size_t bytes);
/* waits until rank `rank` has published `name` with exactly `bytes` bytes; 0, -1 on an I/O error
 * or a size mismatch, -2 when `timeout_s` seconds passed */
int euler_rdv_fetch(const char *dir, const char *name, int rank, void *data, size_t bytes, int timeout_s);

/* Rank 0 passes the communicator id of this run in `uid` (uid_bytes <= 256), the other ranks
 * receive it there.  Every rank gets the same *key.  0, -1 I/O error / bad arguments, -2 timeout
 * (a rank never showed up). */
int euler_rdv_handshake(const char *dir, int rank, int ranks, void *uid, size_t uid_bytes,
                        uint64_t *key, int timeout_s);
/* publish / fetch with the run key in front of the payload: files with another key (or another
 * size) are treated as not yet published */
int euler_rdv_publish_keyed(const char *dir, const char *name, int rank, uint64_t key, const void *data, size_t bytes);
int euler_rdv_fetch_keyed(const char *dir, const char *name, int rank, uint64_t key, void *data, size_t bytes,
                          int timeout_s);
#endif
