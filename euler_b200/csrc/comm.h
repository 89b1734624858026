// euler_b200/csrc/comm.h — row-slab decomposition plumbing: NCCL over
// NVLink/NVSwitch, one process per GPU.  NCCL is bound at run time with dlopen (the library
// has no link-time dependency on it; inside a torch process the already-loaded libnccl is
// reused).  Everything here is enqueued on the handle's stream; nothing synchronises.
#pragma once
#include <stddef.h>

#include "kernels.h"

namespace euler {

struct Comm {
  void* nccl;          // ncclComm_t
  int rank, nranks;
  double* gather;      // device: nranks * GATHER_SLOTS doubles
  unsigned long long* mig;   // device scratch for marker migration counts
  float2 *send_dn, *send_up; // device staging for migrating markers
  size_t send_cap;
};
constexpr int GATHER_SLOTS = 4;

const char* comm_last_error();
int comm_load();                                                  // dlopen + dlsym; 0 ok
int comm_unique_id(void* out128);
int comm_init(Comm* cm, int rank, int nranks, const void* uid128);
void comm_destroy(Comm* cm);

// halo rows of a plane: `depth` rows below own0 and above own1 are refreshed from the
// neighbouring slabs' owned rows (byte-wise, any element size)
int comm_halo(Ctx& c, Comm& cm, void* plane, size_t elem, int depth);
// only the row above own1 (one-directional: p[y+1] of the pressure update)
int comm_halo_up_only(Ctx& c, Comm& cm, void* plane, size_t elem);
// all-gather of GATHER_SLOTS doubles per rank starting at `src` (device) into cm.gather
int comm_gather_scalars(Ctx& c, Comm& cm, const double* src);
int comm_group_begin();
int comm_group_end();
int comm_allreduce_max_u32(Ctx& c, Comm& cm, unsigned int* buf, size_t n);
int comm_allreduce_max_i32(Ctx& c, Comm& cm, int* buf, size_t n);
// point-to-point with both neighbours in one group; pointers may be null when count is 0
int comm_exchange(Ctx& c, Comm& cm, const void* to_dn, size_t n_to_dn, void* from_dn, size_t n_from_dn,
                  const void* to_up, size_t n_to_up, void* from_up, size_t n_from_up);

}  // namespace euler

// ---- peer-to-peer (NVLink) fast path for the per-iteration exchanges ------------------------
// NCCL costs ~40-50 us per small operation; a PCG iteration needs two global scalar exchanges
// and one halo exchange.  With CUDA IPC every rank maps its neighbours' z plane and all ranks'
// mailboxes; partial sums are STORED into every peer's mailbox (double-buffered by sequence
// parity), halo rows are stored straight into the neighbour's halo rows, then a system-scope
// fence and a flag; consumers spin on flags in their OWN memory (with a bounded poll count: a
// lost peer sets DevScalars::comm_timeout instead of hanging the GPU).  Device side: p2p.cuh.
namespace euler {
constexpr int P2P_BLOB_BYTES = 256;      // Mailbox, DistArgs: p2p.cuh

struct P2P {
  bool ready;
  Mailbox* mine;
  Mailbox* peer[P2P_MAX_RANKS];        // peer[r] = rank r's mailbox as mapped here (mine for r == rank)
  double* z_dn;                        // lower / upper neighbour's z plane (row 0 of its storage)
  double* z_up;
  int dn_own1, up_own0;                // their owned-row bounds in their local rows
};

int p2p_export(Ctx& c, Comm& cm, P2P& pp, void* z_raw_base, void* blob /*P2P_BLOB_BYTES*/);
// z_elem: element size of the exchanged z plane (8: fp64 solve, 4: pcg_dtype = FP32)
int p2p_import(Ctx& c, Comm& cm, P2P& pp, const void* blobs /*nranks * P2P_BLOB_BYTES*/, int z_elem = 8);
void p2p_close(P2P& pp, Comm& cm);
// The exchanges as separate small kernels (Ctx::p2p_mode 1; A/B against the fused epilogues):
// stores the top/bottom `depth` owned rows of z into the neighbours' halo rows
void p2p_halo_z(Ctx& c, Comm& cm, P2P& pp, int depth);
// all ranks' DevScalars::part[] -> alpha (kind 0) or beta / sigma / stop test (kind 1);
// wait_halo: also wait for the neighbours' halo rows of this iteration
void p2p_scalars(Ctx& c, Comm& cm, P2P& pp, int kind, bool init, double tol, bool wait_halo);
}  // namespace euler
