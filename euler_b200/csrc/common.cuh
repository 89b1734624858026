// euler_b200/csrc/common.cuh — shared device/host definitions for libeuler_gpu.so (sm_100a).
//
// Data layout in HBM: every plane is row-major [ny][pitch] with x fastest,
// pitch = nx rounded up to 32 elements so each row starts on a 128 B (float) / 256 B
// (double) / 32 B (uint8) boundary, plus GUARD_ROWS rows of zeros before row 0 and after row
// ny-1 so that y±1 / x±1 neighbour loads of border cells stay inside the allocation without
// a bounds branch (the border ring is never fluid: it is all sinks, reference main.c:244-252).
//
// The whole library is compiled with -fmad=false: the reference is built without FMA
// contraction in its parity configuration (oracle/build_ref.sh "strict"), and every fp32
// expression below is written in the reference's evaluation order so results are bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace euler {

constexpr int GUARD_ROWS = 4;
constexpr int PITCH_ALIGN = 32;

enum FaceType { CELL_P = 0, FACE_U = 1, FACE_V = 2 };   // reference celltype_t, main.c:46-50

// A (view of a) grid.  Single GPU: ny == gny, yoff == 0.  Row-slab decomposition: the handle
// stores rows [yoff, yoff+ny) of a grid that is gny rows tall (owned rows plus halo rows), and
// the PCG kernels get a view of the owned rows only (pointers advanced, yoff adjusted).
// Wherever a GLOBAL row matters — sample positions, border clamps, red/black parity, validity
// of the last V row — kernels use y + yoff and gny.
struct Grid {
  int nx, ny;      // columns, rows in this view
  int pitch;       // elements per row
  int yoff;        // global row index of row 0 of this view
  int gny;         // global number of rows
  int th;          // PCG tile height in rows (32, or less when a slab has too few tiles per SM)
};

__host__ __device__ __forceinline__ size_t gidx(const Grid& g, int x, int y) {
  return (size_t)y * (size_t)g.pitch + (size_t)x;
}

// Scalars that live in device memory and are produced/consumed by kernels without a host
// round trip.  One instance per handle.
struct DevScalars {
  // calculate_timestep (main.c:834-841)
  unsigned int max_u2_bits, max_v2_bits;   // float bit patterns (non-negative => uint order)
  float dt, frame_left;
  // markers (main.c:92-95, 204)
  unsigned long long n_markers;
  unsigned long long n_deleted;
  int source_exhausted;
  int pad0;
  unsigned long long rng_state;
  // PCG (main.c:735-767)
  double sigma, zs, alpha, beta, resid;
  double alpha_prev;              // alpha of the iteration before (deferred p update, k_axpy)
  int iters, done, nonzero_rhs, pad1;
  // grid-wide "last block done" counters, one per reducing kernel family
  unsigned int ctr[8];
  // wavefront tickets
  unsigned int ticket[4];
  // faithful marker mode (dt carry-over): number of candidate markers, first fired index
  unsigned long long n_candidates;
  unsigned int active_tiles;
  int marker_overflow;            // more rewinding markers than the candidate list holds
  unsigned long long first_fired;
  // row-slab decomposition: this rank's contributions to the per-iteration all-gather
  // {dot-product partial, ||r||inf partial, source cells needing a marker, marker count}
  double part[4];
  unsigned long long n_send_dn, n_send_up;   // markers leaving towards the lower/upper slab
  unsigned long long src_base, n_markers_global;
  int comm_timeout, pad3;                    // a peer never showed up in a P2P exchange
  unsigned int grid_tiles, pad4;             // tiles the grid stages stream this sub-step (GridTiles)
  // split-phase scalar exchange of the slab solve (p2p.cuh): sigma / alpha by iteration parity (a
  // kernel's block 0 writes the next value while its other blocks still read the current one),
  // and what the host sees of an exchange that has been posted but not consumed yet
  double sigma_s[2], alpha_s[2];
  double peek_resid;
  int peek_done, peek_iters;
  // update_fluid_sources in three parallel passes (marker_kernels.cu): what the emitting pass
  // needs of the state before the bookkeeping pass advanced it
  unsigned long long src_n0, src_state0, src_allow;
  // diagnostics (EULER_TRACE): per-block records of the last traced launch of kernel `trace_kind`
  unsigned long long* trace_blk;
  int trace_kind, pad5;
};

// The MAC-grid stage kernels (grid_kernels.cu) stream only the 512 x 32-cell tiles that can hold a
// nonzero value: tiles with fluid in them or next to them now or in one of the two previous
// sub-steps (everything else is zero in every plane and stays zero).  `list == nullptr` means all
// tx * ty tiles (the first sub-steps after create / reinit / a set() call, and the parity stages).
struct GridTiles {
  const int* list;
  const unsigned int* count;
  int tx, ty;
};
constexpr int GT_W = 512, GT_H = 32;         // tile size in cells
constexpr int GT_SUB = 16;                   // a tile is 4 x 4 thread blocks of 128 cells x 8 rows

// a / h in fp32; the reference's h is 1 (k_side_length, main.c:58) and a / 1.f == a exactly, so
// the common case skips the IEEE division sequence without changing a bit
__device__ __forceinline__ float div_h(float a, float h) { return h == 1.f ? a : a / h; }

#define EULER_FULL_MASK 0xffffffffu

#ifdef __CUDACC__   // device-only from here on (the host build of pcg_ops.cuh stops above)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(EULER_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(EULER_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ float warp_maxf(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(EULER_FULL_MASK, v, o));
  return v;
}

// Deterministic block reduction (blockDim.x*blockDim.y*blockDim.z <= 1024). Result valid in
// thread 0.  IS_MAX selects max instead of sum.
template <bool IS_MAX>
__device__ __forceinline__ double block_reduce(double v) {
  __shared__ double red_smem[32];
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  const int lane = tid & 31, wid = tid >> 5;
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  __syncthreads();                    // protect red_smem reuse across calls
  if (lane == 0) red_smem[wid] = v;
  __syncthreads();
  const int nw = (nthreads + 31) >> 5;
  if (wid == 0) {
    v = lane < nw ? red_smem[lane] : 0.0;
    v = IS_MAX ? warp_max(v) : warp_sum(v);
  }
  return v;
}

// Grid-wide reduction without a second launch and without floating-point atomics: every
// block writes its partial, the block that arrives last re-reduces all partials in a fixed
// order (deterministic for a given grid size) and runs `fin(total)` in its thread 0.
template <bool IS_MAX, class Fin>
__device__ __forceinline__ void grid_reduce_last_block(double block_value, double* partials,
                                                       unsigned int* counter, Fin fin) {
  __shared__ bool is_last;
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (tid == 0) {
    partials[bid] = block_value;
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == nblocks - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc = 0.0;
  for (unsigned int i = tid; i < nblocks; i += nthreads) {
    double p = __ldcg(partials + i);
    acc = IS_MAX ? fmax(acc, p) : acc + p;
  }
  acc = block_reduce<IS_MAX>(acc);
  if (tid == 0) {
    *counter = 0;
    fin(acc);
  }
}

// Same reduction, but control returns to EVERY thread of the block that arrived last (return
// value true, `total` valid in all of its threads), so that the caller can finish the scalar
// step with the whole block — the slab solve exchanges the totals with the other ranks there
// (p2p.cuh).  All other blocks get false.
template <bool IS_MAX>
__device__ __forceinline__ bool grid_reduce_last_block_all(double block_value, double* partials,
                                                           unsigned int* counter, double& total) {
  __shared__ bool is_last_all;
  __shared__ double total_sh;
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
  const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (tid == 0) {
    partials[bid] = block_value;
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    is_last_all = (t == nblocks - 1);
  }
  __syncthreads();
  if (!is_last_all) return false;
  __threadfence();
  double acc = 0.0;
  for (unsigned int i = tid; i < nblocks; i += nthreads) {
    double p = __ldcg(partials + i);
    acc = IS_MAX ? fmax(acc, p) : acc + p;
  }
  acc = block_reduce<IS_MAX>(acc);
  if (tid == 0) {
    *counter = 0;
    total_sh = acc;
  }
  __syncthreads();
  total = total_sh;
  return true;
}
// ---- in-kernel timeline of the two PCG iteration kernels (diagnostics, EULER_TRACE=<slots>) ----
// One 16-word slot per launch; thread 0 of every block folds %globaltimer (ns) into (min, max) pairs:
//   words 0/1 block start, 2/3 scalars known (after the cross-rank collect), 4/5 rows done,
//   6/7 block exit (max = the last block's, after the reduction / post); minima are stored
//   complemented so that one atomicMax serves both.  Word 8 = kernel kind, 9 = iteration.
// A null slot pointer (the default) costs one uniform predicate per mark.
constexpr int TRACE_WORDS = 16;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_mark(unsigned long long* slot, int pair) {
  if (!slot || threadIdx.x != 0) return;
  const unsigned long long t = global_ns();
  atomicMax(slot + 2 * pair, ~t);
  atomicMax(slot + 2 * pair + 1, t);
}
// per-block record of one launch: {start ns, rows-done ns, SM id, 0}
constexpr int TRACE_BLK_WORDS = 4, TRACE_BLK_MAX = 4096;
__device__ __forceinline__ void trace_block(unsigned long long* slot, const DevScalars* sc, int kind, int word) {
  if (!slot || threadIdx.x != 0 || sc->trace_kind != kind || blockIdx.x >= TRACE_BLK_MAX) return;
  unsigned long long* b = sc->trace_blk + (size_t)TRACE_BLK_WORDS * blockIdx.x;
  b[word] = global_ns();
  if (word == 0) { unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); b[2] = smid; }
}
#endif  // __CUDACC__

}  // namespace euler
