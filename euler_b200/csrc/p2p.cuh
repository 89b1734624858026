// euler_b200/csrc/p2p.cuh — device side of the NVLink peer-to-peer exchanges of the slab solve
// ("PCG s halo, every iteration" and "PCG scalars, every iteration").
//
// With CUDA IPC every rank maps its neighbours' z plane and all ranks' mailboxes (comm.cu).
// The exchanges are then plain stores over NVLink followed by a system-scope release of a flag
// in the RECEIVER's memory; consumers spin on flags in their own memory with a bounded poll
// count (a lost peer sets DevScalars::comm_timeout instead of hanging the GPU).
//
// The exchanges are fused into the kernels that produce the data (pcg_kernels.cu):
//   * k_rb_backward_pipe stores the edge rows of z = M^-1 r into the neighbours' halo rows as
//     it computes them, and the block that finishes last raises the neighbours' halo flags,
//     pushes {z.r, ||r||inf} into every rank's mailbox, waits for everybody's and finishes
//     beta / sigma / the stop test (reference main.c:756-765);
//   * k_fused_search_apply's last block does the same for {z.s} -> alpha (main.c:752).
// A PCG iteration on a slab is therefore the same four launches as on a single GPU.
#pragma once
#include "common.cuh"

namespace euler {

constexpr int P2P_MAX_RANKS = 16;
constexpr int P2P_HALO_DEPTH = 4;      // rows of z exchanged per iteration (== SLAB_HALO, api.cu)

struct Mailbox {                       // lives in each rank's device memory, zero-initialised
  double pay[2][P2P_MAX_RANKS][4];     // [sequence parity][sender][slot]
  unsigned long long flag[2][P2P_MAX_RANKS];
  // split-phase exchange: every 8-byte word carries its own flag (high half = sequence tag, low half
  // = 32 bits of payload), so data and flag arrive in ONE atomic store and the sender needs no
  // fence between them (the "LL" idea of NCCL's low-latency protocol)
  unsigned long long ll[2][P2P_MAX_RANKS][4];   // [sequence parity][sender][word]
  unsigned long long halo_flag[2];     // [0] written by the lower neighbour, [1] by the upper
  unsigned long long seq_ctr;          // scalar exchanges this rank has completed (owner-private)
  unsigned long long halo_ctr;         // halo exchanges this rank has posted (owner-private)
  unsigned int halo_done;              // block counter of k_p2p_halo
};

// What a kernel needs to finish a reduction across ranks / to store halo rows into the
// neighbours.  All zero (mine == nullptr) on a single GPU and on the NCCL path.
struct DistArgs {
  Mailbox* mine;
  Mailbox* peer[P2P_MAX_RANKS];        // peer[r] = rank r's mailbox as mapped here
  Mailbox *mb_dn, *mb_up;              // the neighbours' mailboxes (null at the ends)
  double *z_dn, *z_up;                 // neighbours' z plane, biased so that index gidx(view, x, y)
                                       // of an owned edge row lands in the matching halo row
  int rank, nranks;
  int depth;                           // halo rows exchanged
  int pad;
};

constexpr unsigned long long P2P_POLL_LIMIT = 1ull << 24;    // seconds at most; then give up, no hang

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool wait_flag(const unsigned long long* p, unsigned long long want) {
  for (unsigned long long i = 0; i < P2P_POLL_LIMIT; ++i)
    if (ld_acquire_sys(p) >= want) return true;
  return false;
}

// The scalar step that follows a grid-wide reduction, finished across ranks.  Called by ALL
// threads of ONE block (>= 64 threads) after the local reduction is complete and — with
// `halo` — after every block of the grid has fenced its peer stores at system scope.
//   kind 0: part0 = z.s partial            -> alpha                      (main.c:752)
//   kind 1: part0 = z.r partial, part[1] = ||r||inf partial left by k_axpy
//           init: sigma = z.r (main.c:748); else stop test + beta, sigma (main.c:756-765)
// Partials are folded in rank order on every rank: identical, deterministic results.
__device__ __forceinline__ void p2p_finish(const DistArgs& d, DevScalars* sc, int kind, int init,
                                           double tol, double part0, bool halo) {
  const int t = threadIdx.x;
  __shared__ int ok_sh;
  Mailbox* mine = d.mine;
  const unsigned long long seq = mine->seq_ctr;
  const unsigned long long hseq = mine->halo_ctr + 1;
  const int par = (int)(seq & 1ull);
  if (t == 0) ok_sh = 1;
  __syncthreads();
  if (halo) {
    // I am the lower neighbour's UPPER neighbour and vice versa
    if (t == 32 && d.mb_dn) { __threadfence_system(); st_release_sys(&d.mb_dn->halo_flag[1], hseq); }
    if (t == 33 && d.mb_up) { __threadfence_system(); st_release_sys(&d.mb_up->halo_flag[0], hseq); }
  }
  if (t < d.nranks) {
    Mailbox* dst = d.peer[t];
    dst->pay[par][d.rank][0] = part0;
    dst->pay[par][d.rank][1] = sc->part[1];
    __threadfence_system();
    st_release_sys(&dst->flag[par][d.rank], seq + 1);
    if (!wait_flag(&mine->flag[par][t], seq + 1)) ok_sh = 0;
  }
  if (halo) {
    if (t == 34 && d.mb_dn && !wait_flag(&mine->halo_flag[0], hseq)) ok_sh = 0;
    if (t == 35 && d.mb_up && !wait_flag(&mine->halo_flag[1], hseq)) ok_sh = 0;
  }
  __syncthreads();
  if (t != 0) return;
  mine->seq_ctr = seq + 1;
  if (halo) mine->halo_ctr = hseq;
  if (!ok_sh) { sc->comm_timeout = 1; sc->done = 1; return; }
  double sum = 0.0, mx = 0.0;
  for (int r = 0; r < d.nranks; ++r) { sum += mine->pay[par][r][0]; mx = fmax(mx, mine->pay[par][r][1]); }
  if (kind == 0) { sc->zs = sum; sc->alpha_prev = sc->alpha; sc->alpha = sc->sigma / sum; return; }  // main.c:752
  if (init) { sc->sigma = sum; return; }                                      // main.c:748
  sc->resid = mx;
  sc->iters += 1;
  if (mx <= tol) { sc->done = 1; return; }                                    // main.c:756-758
  sc->beta = sum / sc->sigma;                                                 // main.c:762-765
  sc->sigma = sum;
}

// ---- split-phase form of the same exchange (EULER_P2P_SPLIT=1) ---------------------------------
// p2p_finish makes the PRODUCING kernel's last block wait for every rank's partial: an NVLink
// round trip plus the skew between ranks, serialised with the kernel's tail and the next launch.
// Here the last block only POSTS (stores + release, no wait) and every block of the NEXT kernel
// COLLECTS at its start: the wait overlaps the launch gap and the next kernel's ramp, and the
// scalar step (alpha, or the stop test and beta) is repeated identically by every block.
// Reuse of a mailbox slot (parity of the sequence number) is safe for the same reason as before:
// a peer posts exchange n+2 only after it collected n+1, which needs my post of n+1, which my
// last block makes after all my blocks collected n.
// Not inlined on purpose: with the exchange code inlined, the register allocation of the fused tail
// kernel's row loop changed with the protocol (measured on B200: 663 vs 652 us per even iteration at
// 16384^2 for code that never runs on one GPU, profiles/r02e_trace_*); behind a call the row loop
// compiles the same whatever the exchange looks like.  One call per kernel.
#ifdef EULER_P2P_INLINE
#define P2P_LL_INLINE __forceinline__
#else
#define P2P_LL_INLINE __noinline__
#endif
#ifndef EULER_P2P_LL
#define EULER_P2P_LL 3      // A/B builds: bit 0 = {z.s} (search -> tail), bit 1 = {z.r, ||r||inf} + halo (tail -> search)
#endif
// sequence tag of exchange `seq` (never 0: a zero-initialised mailbox holds no message)
__device__ __forceinline__ unsigned int ll_tag(unsigned long long seq) {
  return (unsigned int)(seq % 0xfffffffeull) + 1u;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// All threads of the block that finished the local reduction.  `halo`: this kernel's blocks stored
// edge rows of z into the neighbours (each storing thread fenced at system scope before its block
// took its ticket): the message to a NEIGHBOUR doubles as its halo flag, so the posting thread
// orders it behind those stores with one system-scope fence (cumulative over everything that
// happened before it on this GPU).  Messages to the other ranks leave at once.
static __device__ P2P_LL_INLINE void p2p_post_ll(const DistArgs& d, double part0, double part1, bool halo) {
  const int t = threadIdx.x;
  Mailbox* mine = d.mine;
  const unsigned long long seq = mine->seq_ctr;
  const int par = (int)(seq & 1ull);
  if (t < d.nranks) {
    Mailbox* dst = d.peer[t];
    if (halo && (dst == d.mb_dn || dst == d.mb_up)) __threadfence_system();
    const unsigned long long tag = (unsigned long long)ll_tag(seq) << 32;
    const unsigned long long a = (unsigned long long)__double_as_longlong(part0);
    const unsigned long long b = (unsigned long long)__double_as_longlong(part1);
    unsigned long long* w = dst->ll[par][d.rank];
    st_relaxed_sys(w + 0, tag | (a & 0xffffffffull));
    st_relaxed_sys(w + 1, tag | (a >> 32));
    st_relaxed_sys(w + 2, tag | (b & 0xffffffffull));
    st_relaxed_sys(w + 3, tag | (b >> 32));
  }
  __syncthreads();
  if (t == 0) mine->seq_ctr = seq + 1;
}

// All threads of a block (>= 64).  false: a peer never showed up (bounded poll).  `wait_halo`: the
// neighbours' messages also announce their halo rows of z; the polling threads acquire at system
// scope before the block goes on to read them.
static __device__ P2P_LL_INLINE bool p2p_collect_ll(const DistArgs& d, bool wait_halo, double& sum, double& mx) {
  const int t = threadIdx.x;
  __shared__ int ok_c;
  __shared__ unsigned int half_c[P2P_MAX_RANKS][4];
  __shared__ double res_c[2];
  Mailbox* mine = d.mine;
  const unsigned long long seq = mine->seq_ctr;             // posted by my own rank's previous kernel
  const int par = (int)((seq - 1ull) & 1ull);
  const unsigned int tag = ll_tag(seq - 1ull);
  if (t == 0) ok_c = 1;
  __syncthreads();
  if (t < 4 * d.nranks) {                                    // one thread per (sender, word)
    const unsigned long long* w = &mine->ll[par][t >> 2][t & 3];
    unsigned long long v = 0;
    bool got = false;
    for (unsigned long long i = 0; i < P2P_POLL_LIMIT && !got; ++i) {
      v = ld_relaxed_sys(w);
      got = (unsigned int)(v >> 32) == tag;
    }
    if (!got) ok_c = 0;
    half_c[t >> 2][t & 3] = (unsigned int)v;
#ifndef EULER_P2P_NOFENCE
    if (wait_halo) __threadfence_system();
#endif
  }
  __syncthreads();
  if (t == 0) {
    double a = 0.0, m = 0.0;
    for (int r = 0; r < d.nranks; ++r) {                     // rank order on every rank: identical results
      const double p0 = __longlong_as_double((long long)(((unsigned long long)half_c[r][1] << 32) | half_c[r][0]));
      const double p1 = __longlong_as_double((long long)(((unsigned long long)half_c[r][3] << 32) | half_c[r][2]));
      a += p0; m = fmax(m, p1);
    }
    res_c[0] = a; res_c[1] = m;
  }
  __syncthreads();
  sum = res_c[0]; mx = res_c[1];
  return ok_c != 0;
}

// ---- the flag-per-sender form with release stores (round-2 first version; A/B builds) ----
__device__ __forceinline__ void p2p_post_fl(const DistArgs& d, double part0, double part1, bool halo) {
  const int t = threadIdx.x;
  Mailbox* mine = d.mine;
  const unsigned long long seq = mine->seq_ctr;
  const unsigned long long hseq = mine->halo_ctr + 1;
  const int par = (int)(seq & 1ull);
  if (halo) {
    if (t == 32 && d.mb_dn) { __threadfence_system(); st_release_sys(&d.mb_dn->halo_flag[1], hseq); }
    if (t == 33 && d.mb_up) { __threadfence_system(); st_release_sys(&d.mb_up->halo_flag[0], hseq); }
  }
  if (t < d.nranks) {
    Mailbox* dst = d.peer[t];
    dst->pay[par][d.rank][0] = part0;
    dst->pay[par][d.rank][1] = part1;
    st_release_sys(&dst->flag[par][d.rank], seq + 1);       // release orders the two stores before it
  }
  __syncthreads();
  if (t == 0) { mine->seq_ctr = seq + 1; if (halo) mine->halo_ctr = hseq; }
}

// All threads of a block (>= 64).  false: a peer never showed up (bounded poll).
__device__ __forceinline__ bool p2p_collect_fl(const DistArgs& d, bool wait_halo, double& sum, double& mx) {
  const int t = threadIdx.x;
  __shared__ int ok_c;
  __shared__ double res_c[2];
  Mailbox* mine = d.mine;
  const unsigned long long seq = mine->seq_ctr;             // posted by my own rank's previous kernel
  const int par = (int)((seq - 1ull) & 1ull);
  if (t == 0) ok_c = 1;
  __syncthreads();
  if (t < d.nranks && !wait_flag(&mine->flag[par][t], seq)) ok_c = 0;
  if (wait_halo) {
    const unsigned long long hseq = mine->halo_ctr;
    if (t == 34 && d.mb_dn && !wait_flag(&mine->halo_flag[0], hseq)) ok_c = 0;
    if (t == 35 && d.mb_up && !wait_flag(&mine->halo_flag[1], hseq)) ok_c = 0;
  }
  __syncthreads();
  if (t == 0) {
    double a = 0.0, m = 0.0;
    for (int r = 0; r < d.nranks; ++r) { a += mine->pay[par][r][0]; m = fmax(m, mine->pay[par][r][1]); }
    res_c[0] = a; res_c[1] = m;
  }
  __syncthreads();
  sum = res_c[0]; mx = res_c[1];
  return ok_c != 0;
}


__device__ __forceinline__ void p2p_post(const DistArgs& d, double part0, double part1, bool halo) {
  if ((halo ? (EULER_P2P_LL & 2) : (EULER_P2P_LL & 1)) != 0) p2p_post_ll(d, part0, part1, halo);
  else p2p_post_fl(d, part0, part1, halo);
}
__device__ __forceinline__ bool p2p_collect(const DistArgs& d, bool wait_halo, double& sum, double& mx) {
  if ((wait_halo ? (EULER_P2P_LL & 2) : (EULER_P2P_LL & 1)) != 0) return p2p_collect_ll(d, wait_halo, sum, mx);
  return p2p_collect_fl(d, wait_halo, sum, mx);
}

}  // namespace euler
