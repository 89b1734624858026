/* euler_b200/host/rendezvous.h — file rendezvous between the N processes of one slab-decomposed
 * run (one process per GPU).  The C-ABI leaves the side channel to the caller
 * (euler_gpu_comm_init takes the 128-byte communicator id "distributed over any side channel");
 * bench.py uses torch.distributed for it, the host C program uses a directory the ranks share:
 * rank r publishes <dir>/<name><r>.bin atomically (write to a temporary name, then rename) and
 * the others poll for it.  The directory must be fresh for every run. */
#ifndef EULER_RENDEZVOUS_H
#define EULER_RENDEZVOUS_H
#include <stddef.h>

/* 0, or -1 on an I/O error */
int euler_rdv_publish(const char *dir, const char *name, int rank, const void *data, size_t bytes);
/* waits until rank `rank` has published `name` with exactly `bytes` bytes; 0, -1 on an I/O error
 * or a size mismatch, -2 when `timeout_s` seconds passed */
int euler_rdv_fetch(const char *dir, const char *name, int rank, void *data, size_t bytes, int timeout_s);
#endif
