// euler_b200/csrc/comm.cu — see comm.h.
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace euler {

namespace {

thread_local char g_cerr[256] = "";

struct Api {
  void* dl;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
} api = {};

int nerr(ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return 0;
  snprintf(g_cerr, sizeof g_cerr, "%s: %s", what, api.GetErrorString ? api.GetErrorString(r) : "nccl error");
  return -1;
}
#define NC(call) do { if (nerr((call), #call)) return -1; } while (0)

}  // namespace

const char* comm_last_error() { return g_cerr; }

int comm_load() {
  if (api.dl) return 0;
  void* dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // torch's, if loaded
  if (!dl) dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!dl) dl = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!dl) { snprintf(g_cerr, sizeof g_cerr, "cannot load libnccl.so.2: %s", dlerror()); return -1; }
#define SYM(field, name) do { *(void**)(&api.field) = dlsym(dl, name); \
    if (!api.field) { snprintf(g_cerr, sizeof g_cerr, "libnccl lacks %s", name); return -1; } } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy"); SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  api.dl = dl;
  return 0;
}

int comm_unique_id(void* out128) {
  if (comm_load()) return -1;
  ncclUniqueId id;
  NC(api.GetUniqueId(&id));
  memcpy(out128, &id, sizeof id);
  return 0;
}

int comm_init(Comm* cm, int rank, int nranks, const void* uid128) {
  if (comm_load()) return -1;
  ncclUniqueId id;
  memcpy(&id, uid128, sizeof id);
  ncclComm_t comm = nullptr;
  NC(api.CommInitRank(&comm, nranks, id, rank));
  cm->nccl = comm; cm->rank = rank; cm->nranks = nranks;
  return 0;
}

void comm_destroy(Comm* cm) {
  if (cm && cm->nccl && api.CommDestroy) api.CommDestroy((ncclComm_t)cm->nccl);
  if (cm) cm->nccl = nullptr;
}

int comm_exchange(Ctx& c, Comm& cm, const void* to_dn, size_t n_to_dn, void* from_dn, size_t n_from_dn,
                  const void* to_up, size_t n_to_up, void* from_up, size_t n_from_up) {
  ncclComm_t comm = (ncclComm_t)cm.nccl;
  const int dn = cm.rank - 1, up = cm.rank + 1;
  NC(api.GroupStart());
  // a failed call must not leave the group open: remember the first error, always close
  int bad = 0;
#define NCG(call) do { if (!bad && nerr((call), #call)) bad = 1; } while (0)
  if (dn >= 0) {
    if (n_to_dn) NCG(api.Send(to_dn, n_to_dn, ncclChar, dn, comm, c.stream));
    if (n_from_dn) NCG(api.Recv(from_dn, n_from_dn, ncclChar, dn, comm, c.stream));
  }
  if (up < cm.nranks) {
    if (n_to_up) NCG(api.Send(to_up, n_to_up, ncclChar, up, comm, c.stream));
    if (n_from_up) NCG(api.Recv(from_up, n_from_up, ncclChar, up, comm, c.stream));
  }
#undef NCG
  const ncclResult_t end = api.GroupEnd();
  if (bad) return -1;
  NC(end);
  return 0;
}

int comm_halo(Ctx& c, Comm& cm, void* plane, size_t elem, int depth) {
  char* base = reinterpret_cast<char*>(plane);
  const size_t row = (size_t)c.g.pitch * elem;
  const int d_dn = depth < c.own0 ? depth : c.own0;                  // halo rows I store below
  const int d_up = depth < c.g.ny - c.own1 ? depth : c.g.ny - c.own1;
  // what I send is what the neighbour stores as halo: the same depth (slabs are >= depth rows)
  return comm_exchange(c, cm,
                       base + (size_t)c.own0 * row, cm.rank > 0 ? (size_t)depth * row : 0,
                       base + (size_t)(c.own0 - d_dn) * row, (size_t)d_dn * row,
                       base + (size_t)(c.own1 - depth) * row, cm.rank + 1 < cm.nranks ? (size_t)depth * row : 0,
                       base + (size_t)c.own1 * row, (size_t)d_up * row);
}

int comm_halo_up_only(Ctx& c, Comm& cm, void* plane, size_t elem) {
  char* base = reinterpret_cast<char*>(plane);
  const size_t row = (size_t)c.g.pitch * elem;
  const bool has_up = cm.rank + 1 < cm.nranks && c.g.ny > c.own1;
  return comm_exchange(c, cm, base + (size_t)c.own0 * row, cm.rank > 0 ? row : 0, nullptr, 0,
                       nullptr, 0, base + (size_t)c.own1 * row, has_up ? row : 0);
}

// NCCL groups nest: operations issued between begin/end (including the groups comm_exchange
// opens itself) are fused into one launch
int comm_group_begin() { if (comm_load()) return -1; NC(api.GroupStart()); return 0; }
int comm_group_end() { NC(api.GroupEnd()); return 0; }

int comm_gather_scalars(Ctx& c, Comm& cm, const double* src) {
  NC(api.AllGather(src, cm.gather, GATHER_SLOTS, ncclDouble, (ncclComm_t)cm.nccl, c.stream));
  return 0;
}

int comm_allreduce_max_u32(Ctx& c, Comm& cm, unsigned int* buf, size_t n) {
  NC(api.AllReduce(buf, buf, n, ncclUint32, ncclMax, (ncclComm_t)cm.nccl, c.stream));
  return 0;
}
int comm_allreduce_max_i32(Ctx& c, Comm& cm, int* buf, size_t n) {
  NC(api.AllReduce(buf, buf, n, ncclInt32, ncclMax, (ncclComm_t)cm.nccl, c.stream));
  return 0;
}

}  // namespace euler

// =============================================================================== P2P ====
namespace euler {

namespace {

struct Blob {                          // what every rank tells the others (fits P2P_BLOB_BYTES)
  cudaIpcMemHandle_t z, mbox;
  // cudaMalloc sub-allocates small buffers inside larger driver allocations and an IPC handle
  // always maps the WHOLE allocation: byte offsets of our buffers inside theirs
  unsigned long long z_off, mbox_off;
  int own0, own1, ny, pitch;
};

// base address of the driver allocation that contains p (cuMemGetAddressRange, bound through
// the runtime so that libcuda is not a link-time dependency)
int alloc_base(const void* p, unsigned long long* off) {
  typedef int (*fn_t)(unsigned long long*, size_t*, unsigned long long);
  static fn_t fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) {
      snprintf(g_cerr, sizeof g_cerr, "cuMemGetAddressRange not available");
      return -1;
    }
    fn = (fn_t)sym;
  }
  unsigned long long base = 0;
  size_t size = 0;
  if (fn(&base, &size, (unsigned long long)(uintptr_t)p) != 0) {
    snprintf(g_cerr, sizeof g_cerr, "cuMemGetAddressRange failed");
    return -1;
  }
  *off = (unsigned long long)(uintptr_t)p - base;
  return 0;
}
static_assert(sizeof(Blob) <= P2P_BLOB_BYTES, "blob too large");

struct PeerPtrs { Mailbox* p[P2P_MAX_RANKS]; };

// One block.  Thread t < nranks: store my partials into rank t's mailbox, then wait for rank
// t's partials in mine.  Thread 0 folds them in rank order (deterministic) and finishes the
// scalar step exactly like k_dist_alpha / k_dist_beta.  (Separate-kernel form of p2p_finish.)
__global__ void __launch_bounds__(32) k_p2p_scalars(PeerPtrs peers, Mailbox* mine, DevScalars* sc,
                                                    int rank, int nranks, int kind, int init,
                                                    double tol, int wait_halo, int has_dn, int has_up) {
  const int t = threadIdx.x;
  const unsigned long long seq = mine->seq_ctr, halo_seq = mine->halo_ctr;
  const int par = (int)(seq & 1ull);
  __shared__ int ok_sh;
  if (t == 0) ok_sh = 1;
  __syncthreads();
  if (t < nranks) {
    Mailbox* dst = peers.p[t];
#pragma unroll
    for (int k = 0; k < 4; ++k) dst->pay[par][rank][k] = sc->part[k];
    __threadfence_system();
    st_release_sys(&dst->flag[par][rank], seq + 1);
    if (!wait_flag(&mine->flag[par][t], seq + 1)) ok_sh = 0;
  }
  if (wait_halo) {
    if (t == 30 && has_dn && !wait_flag(&mine->halo_flag[0], halo_seq)) ok_sh = 0;
    if (t == 31 && has_up && !wait_flag(&mine->halo_flag[1], halo_seq)) ok_sh = 0;
  }
  __syncthreads();
  if (t != 0) return;
  mine->seq_ctr = seq + 1;
  if (!ok_sh) { sc->comm_timeout = 1; sc->done = 1; return; }
  if (sc->done) return;
  double sum = 0.0, mx = 0.0;
  for (int r = 0; r < nranks; ++r) { sum += mine->pay[par][r][0]; mx = fmax(mx, mine->pay[par][r][1]); }
  if (kind == 0) { sc->zs = sum; sc->alpha_prev = sc->alpha; sc->alpha = sc->sigma / sum; return; }  // main.c:752
  if (init) { sc->sigma = sum; return; }                                      // main.c:748
  sc->resid = mx;
  sc->iters += 1;
  if (mx <= tol) { sc->done = 1; return; }                                    // main.c:756-758
  sc->beta = sum / sc->sigma;                                                 // main.c:762-765
  sc->sigma = sum;
}

// Stores my edge rows straight into the neighbours' halo rows over NVLink (16 B stores), then
// the block that finishes last fences and raises the neighbours' flags.
__global__ void __launch_bounds__(256) k_p2p_halo(const double* __restrict__ z, int pitch, int own0,
                                                  int own1, int depth, double* z_dn, int dn_own1,
                                                  double* z_up, int up_own0, Mailbox* mine,
                                                  Mailbox* mb_dn, Mailbox* mb_up) {
  const size_t n2 = (size_t)depth * pitch / 2;             // double2 elements per direction
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += stride) {
    if (z_dn) {      // my lowest owned rows -> the halo rows just above the lower neighbour's owned rows
      const double2 v = reinterpret_cast<const double2*>(z + (size_t)own0 * pitch)[i];
      reinterpret_cast<double2*>(z_dn + (size_t)dn_own1 * pitch)[i] = v;
    }
    if (z_up) {      // my highest owned rows -> the halo rows just below the upper neighbour's owned rows
      const double2 v = reinterpret_cast<const double2*>(z + (size_t)(own1 - depth) * pitch)[i];
      reinterpret_cast<double2*>(z_up + (size_t)(up_own0 - depth) * pitch)[i] = v;
    }
  }
  __threadfence_system();
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&mine->halo_done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  mine->halo_done = 0;
  const unsigned long long halo_seq = mine->halo_ctr + 1;
  mine->halo_ctr = halo_seq;
  __threadfence_system();
  if (mb_dn) st_release_sys(&mb_dn->halo_flag[1], halo_seq);   // I am its UPPER neighbour
  if (mb_up) st_release_sys(&mb_up->halo_flag[0], halo_seq);   // I am its LOWER neighbour
}

}  // namespace

#define CUQ(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(g_cerr, sizeof g_cerr, "%s: %s", #call, cudaGetErrorString(e_)); return -1; } } while (0)

int p2p_export(Ctx& c, Comm& cm, P2P& pp, void* z_raw_base, void* blob) {
  (void)cm;
  if (!pp.mine) {
    CUQ(cudaMalloc((void**)&pp.mine, sizeof(Mailbox)));
    CUQ(cudaMemset(pp.mine, 0, sizeof(Mailbox)));
  }
  Blob b;
  memset(&b, 0, sizeof b);
  CUQ(cudaIpcGetMemHandle(&b.z, z_raw_base));
  CUQ(cudaIpcGetMemHandle(&b.mbox, pp.mine));
  if (alloc_base(z_raw_base, &b.z_off) || alloc_base(pp.mine, &b.mbox_off)) return -1;
  b.own0 = c.own0; b.own1 = c.own1; b.ny = c.g.ny; b.pitch = c.g.pitch;
  memset(blob, 0, P2P_BLOB_BYTES);
  memcpy(blob, &b, sizeof b);
  return 0;
}

int p2p_import(Ctx& c, Comm& cm, P2P& pp, const void* blobs, int z_elem) {
  if (cm.nranks > P2P_MAX_RANKS) { snprintf(g_cerr, sizeof g_cerr, "p2p: more than %d ranks", P2P_MAX_RANKS); return -1; }
  const char* base = reinterpret_cast<const char*>(blobs);
  for (int r = 0; r < cm.nranks; ++r) {
    Blob b;
    memcpy(&b, base + (size_t)r * P2P_BLOB_BYTES, sizeof b);
    if (b.pitch != c.g.pitch) { snprintf(g_cerr, sizeof g_cerr, "p2p: pitch mismatch"); return -1; }
    if (r == cm.rank) { pp.peer[r] = pp.mine; continue; }
    void* p = nullptr;
    CUQ(cudaIpcOpenMemHandle(&p, b.mbox, cudaIpcMemLazyEnablePeerAccess));
    pp.peer[r] = reinterpret_cast<Mailbox*>(reinterpret_cast<char*>(p) + b.mbox_off);
    if (r == cm.rank - 1 || r == cm.rank + 1) {
      void* zp = nullptr;
      CUQ(cudaIpcOpenMemHandle(&zp, b.z, cudaIpcMemLazyEnablePeerAccess));
      // (typed double* whatever the element size: the launchers bias it in bytes)
      double* row0 = reinterpret_cast<double*>(reinterpret_cast<char*>(zp) + b.z_off +
                                               (size_t)GUARD_ROWS * c.g.pitch * (size_t)z_elem);
      if (r == cm.rank - 1) { pp.z_dn = row0; pp.dn_own1 = b.own1; }
      else { pp.z_up = row0; pp.up_own0 = b.own0; }
    }
  }
  // what the kernels need (p2p.cuh); z_dn / z_up are biased per view by the launchers
  DistArgs& d = c.dist;
  memset(&d, 0, sizeof d);
  d.mine = pp.mine;
  for (int r = 0; r < cm.nranks; ++r) d.peer[r] = pp.peer[r];
  d.mb_dn = cm.rank > 0 ? pp.peer[cm.rank - 1] : nullptr;
  d.mb_up = cm.rank + 1 < cm.nranks ? pp.peer[cm.rank + 1] : nullptr;
  d.z_dn = pp.z_dn; d.z_up = pp.z_up;
  d.rank = cm.rank; d.nranks = cm.nranks;
  c.p2p_dn_own1 = pp.dn_own1; c.p2p_up_own0 = pp.up_own0;
  pp.ready = true;
  return 0;
}

void p2p_close(P2P& pp, Comm& cm) {
  if (!pp.mine) return;
  // peer mappings are released with the process (their base addresses are not kept)
  (void)cm;
  cudaFree(pp.mine);
  pp.mine = nullptr; pp.ready = false;
}

void p2p_halo_z(Ctx& c, Comm& cm, P2P& pp, int depth) {
  Mailbox* mb_dn = cm.rank > 0 ? pp.peer[cm.rank - 1] : nullptr;
  Mailbox* mb_up = cm.rank + 1 < cm.nranks ? pp.peer[cm.rank + 1] : nullptr;
  k_p2p_halo<<<32, 256, 0, c.stream>>>(c.z, c.g.pitch, c.own0, c.own1, depth, pp.z_dn, pp.dn_own1,
                                       pp.z_up, pp.up_own0, pp.mine, mb_dn, mb_up);
  c.launches += 1;
}

void p2p_scalars(Ctx& c, Comm& cm, P2P& pp, int kind, bool init, double tol, bool wait_halo) {
  PeerPtrs peers;
  for (int r = 0; r < P2P_MAX_RANKS; ++r) peers.p[r] = r < cm.nranks ? pp.peer[r] : nullptr;
  k_p2p_scalars<<<1, 32, 0, c.stream>>>(peers, pp.mine, c.sc, cm.rank, cm.nranks, kind,
                                        init ? 1 : 0, tol, wait_halo ? 1 : 0,
                                        cm.rank > 0 ? 1 : 0, cm.rank + 1 < cm.nranks ? 1 : 0);
  c.launches += 1;
}

}  // namespace euler
