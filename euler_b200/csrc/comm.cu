// euler_b200/csrc/comm.cu — see comm.h.
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>

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
  if (dn >= 0) {
    if (n_to_dn) NC(api.Send(to_dn, n_to_dn, ncclChar, dn, comm, c.stream));
    if (n_from_dn) NC(api.Recv(from_dn, n_from_dn, ncclChar, dn, comm, c.stream));
  }
  if (up < cm.nranks) {
    if (n_to_up) NC(api.Send(to_up, n_to_up, ncclChar, up, comm, c.stream));
    if (n_from_up) NC(api.Recv(from_up, n_from_up, ncclChar, up, comm, c.stream));
  }
  NC(api.GroupEnd());
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
