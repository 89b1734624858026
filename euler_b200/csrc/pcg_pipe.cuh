// euler_b200/csrc/pcg_pipe.cuh — TMA-fed row pipeline for the 5-point stencil kernels of the
// pressure solve (apply_a, red-black forward/backward).
//
// Why: with a register window the loads of the next row are consumed almost immediately, so
// the bytes in flight are bounded by registers x occupancy and the kernels sat at ~45 % of
// HBM bandwidth (profiles/r01a_*).  Here the bulk-copy engine keeps the memory system busy
// independently of the SM's registers:
//
//   * a persistent block walks its 512 x TH tiles row by row; every input-plane row segment of
//     a tile (with its halo columns) is ONE contiguous, 16 B-aligned range, fetched with
//     `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` (SASS: UBLKCP) into a
//     ring of NS shared-memory stages;
//   * `full[stage]` mbarriers carry the transaction bytes; `empty[stage]` mbarriers are
//     arrived on by the 4 consumer warps when a row is no longer needed (a row lives for
//     three output rows: as up-, centre- and down-neighbour);
//   * one elected thread issues the copies NS-3 rows ahead of the row being computed and keeps
//     going across tile boundaries, so the pipeline never drains inside a launch;
//   * consumers read centre/left/right/up/down straight from shared memory (no shuffles, no
//     edge-lane special case) and store results with 16 B vector stores.
#pragma once
#include "common.cuh"

namespace euler {
namespace pipe {

constexpr int TW = 512, TT = 128;            // tile width in cells, threads per block
constexpr int HX1 = 16;                      // halo columns of a u8 plane (16 B)
constexpr int ROW1 = (TW + 2 * HX1);         // bytes of one u8 row segment in smem
// The value planes are fp64 (the reference's precision) or fp32 (mixed-precision mode,
// pcg_dtype = FP32): T is the element type, a row segment carries 16 B of halo columns on
// each side whatever T is (bulk copies need 16 B-aligned addresses and sizes).
template <class T> struct Elem {
  static constexpr int HX = 16 / (int)sizeof(T);          // halo columns: 2 (fp64) / 4 (fp32)
  static constexpr int ROW = (TW + 2 * HX) * (int)sizeof(T);   // bytes of one row segment in smem
};

#ifdef __CUDACC__   // mbarrier / bulk-copy PTX: device compilation only
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  while (!mbar_try_wait(b, parity)) {}
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// L2 prefetch of a contiguous global range by the bulk-copy engine (no destination, no mbarrier):
// for operands that are read with plain loads a few rows later
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(src)), "r"(bytes)
               : "memory");
}

#endif  // __CUDACC__

struct Tiles { int tx, ty, n; };
__host__ __device__ inline Tiles tiles_of(const Grid& g, int th) {
  Tiles t;
  t.tx = (g.pitch + TW - 1) / TW;
  t.ty = (g.ny + th - 1) / th;
  t.n = t.tx * t.ty;
  return t;
}

// Work split of the persistent kernels over the ordered list of active tiles (n tiles, B blocks,
// n = q B + rem):
//   * the first q B tiles go round robin, whole tiles: block b takes tiles b, b + B, ...  All
//     blocks advance at the same pace, so at any moment the grid streams ONE contiguous window
//     of B tiles — good DRAM page locality (a contiguous private range per block measured
//     5 % slower on the pure streaming kernel, profiles/r01e);
//   * the last `rem` tiles are cut by ROWS: seen as one long column of rem x th "strip rows",
//     block b owns the contiguous range [b R / B, (b+1) R / B) of it.  Every block gets the
//     same number of rows to within one, instead of `rem` blocks working a whole tile longer
//     than the others (10 % of the kernel at 4.5 tiles per block, 40 % at 1.5; on a thin slab
//     or a 4096^2 grid n < B and everything is split this way).
// A row range is cut into PIECES at tile boundaries; each piece is a run of consecutive rows
// of one tile and needs its own halo rows.
struct Piece { int x0, y0, y1, w; };

#ifdef __CUDACC__   // the work split reads blockIdx / gridDim
struct ChunkIter {
  int k, q;                        // next round robin round, number of full rounds
  int cur, end;                    // row-split tail, in strip rows from the start of the list;
                                   // n * th < 2^31 (grids are < 2^30 cells)
  __device__ __forceinline__ void start(int n_active, int th) {
    const int B = (int)gridDim.x;
    q = n_active / B;
    k = 0;
    const long R = (long)(n_active - q * B) * th, base = (long)q * B * th;
    cur = (int)(base + R * blockIdx.x / B);
    end = (int)(base + R * (blockIdx.x + 1) / B);
  }
  __device__ __forceinline__ bool next(const Grid& g, const Tiles& T, int th,
                                       const int* __restrict__ list, Piece& p) {
    for (;;) {
      int tile, r0, take;
      if (k < q) {
        tile = list[blockIdx.x + k * (int)gridDim.x];
        ++k;
        r0 = 0; take = th;
      } else if (cur < end) {
        const int pos = cur / th;
        r0 = cur - pos * th;
        take = min(end - cur, th - r0);
        tile = list[pos];
        cur += take;
      } else {
        return false;
      }
      const int ty = tile / T.tx;
      p.y0 = ty * th + r0;
      p.y1 = min(p.y0 + take, g.ny);           // the top row of tiles may be cut by the grid
      if (p.y0 < p.y1) {
        p.x0 = (tile - ty * T.tx) * TW;
        p.w = min(TW, g.pitch - p.x0);
        return true;
      }
    }
  }
};

// Enumerates the rows (jobs) a block has to load: for every piece of its range, rows
// y0-rad .. y1-1+rad inclusive.
struct JobIter {
  ChunkIter ch;
  Piece p;
  int yy;
  int rad = 1;                     // halo rows loaded below/above the piece
  bool valid;
  __device__ __forceinline__ void seek(const Grid& g, const Tiles& T, int th,
                                       const int* __restrict__ list) {
    valid = ch.next(g, T, th, list, p);
    if (valid) yy = p.y0 - rad;
  }
  __device__ __forceinline__ void start(const Grid& g, const Tiles& T, int th,
                                        const int* __restrict__ list, int n) {
    ch.start(n, th);
    seek(g, T, th, list);
  }
  __device__ __forceinline__ void next(const Grid& g, const Tiles& T, int th,
                                       const int* __restrict__ list) {
    if (yy < p.y1 - 1 + rad) { ++yy; return; }
    seek(g, T, th, list);
  }
};

#endif  // __CUDACC__

// One stage of the ring, as seen by the consumers: pointers are biased so that index i is tile
// column i (value planes valid for i in [-HX, TW+HX), u8 planes for i in [-16, TW+16)).
template <int ND, int NB, class T = double>
struct RowView {
  const T* d[ND];
  const uint8_t* b[NB];
};

template <int ND, int NB, class T = double>
struct Layout {
  static constexpr int ROWT = Elem<T>::ROW, HXT = Elem<T>::HX;
  static constexpr int stage_bytes = ND * ROWT + NB * ROW1;
  __device__ static __forceinline__ RowView<ND, NB, T> view(unsigned char* stage) {
    RowView<ND, NB, T> v;
#pragma unroll
    for (int i = 0; i < ND; ++i) v.d[i] = reinterpret_cast<const T*>(stage + i * ROWT) + HXT;
#pragma unroll
    for (int i = 0; i < NB; ++i) v.b[i] = stage + ND * ROWT + i * ROW1 + HX1;
    return v;
  }
};

template <int ND, int NB, class T = double>
struct Planes {
  const T* d[ND];
  const uint8_t* b[NB];
};

#ifdef __CUDACC__
// The pipeline driver.  `Op::row(dn, ce, up, t4, x, y, live)` is called by every thread for
// each output row of each active tile of the block: t4 = 4*threadIdx.x is the tile column of
// the thread's first cell, x its global column; `live` is false for threads past the row end.
template <int ND, int NB, int NS, int TH, class Op, int C = 4, class E = double>
__device__ __forceinline__ void run(const Grid& g, const int* __restrict__ list, int n_active,
                                    const Planes<ND, NB, E>& in, Op& op) {
  using L = Layout<ND, NB, E>;
  constexpr int ROWT = L::ROWT, HXT = L::HXT;
  constexpr int PF = NS - 3;
  constexpr int NTHREADS = TW / C;                 // C cells per thread
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* stages = smem_raw;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + NS * L::stage_bytes);
  uint64_t* empty = full + NS;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, NTHREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int th = g.th;                             // run-time tile height (TH is the default)
  const Tiles T = tiles_of(g, th);
  JobIter cons, prod;
  cons.start(g, T, th, list, n_active);
  prod = cons;
  int issued = 0;

  auto issue = [&]() {          // thread 0 only
    const int s = issued % NS, use = issued / NS;
    mbar_wait(empty + s, (use & 1) ^ 1);
    unsigned char* st = stages + s * L::stage_bytes;
    const uint32_t b8 = (uint32_t)(prod.p.w + 2 * HXT) * (uint32_t)sizeof(E), b1 = (uint32_t)(prod.p.w + 2 * HX1);
    mbar_expect_tx(full + s, ND * b8 + NB * b1);
    const long row = (long)prod.yy * g.pitch + prod.p.x0;
#pragma unroll
    for (int i = 0; i < ND; ++i) bulk_g2s(st + i * ROWT, in.d[i] + row - HXT, b8, full + s);
#pragma unroll
    for (int i = 0; i < NB; ++i) bulk_g2s(st + ND * ROWT + i * ROW1, in.b[i] + row - HX1, b1, full + s);
    ++issued;
    prod.next(g, T, th, list);
  };

  for (int j = 0; cons.valid; ++j) {
    if (threadIdx.x == 0)
      while (prod.valid && issued <= j + PF) issue();
    mbar_wait(full + (j % NS), (j / NS) & 1);
    if (cons.yy >= cons.p.y0 + 1) {
      const RowView<ND, NB, E> up = L::view(stages + (j % NS) * L::stage_bytes);
      const RowView<ND, NB, E> ce = L::view(stages + ((j + NS - 1) % NS) * L::stage_bytes);
      const RowView<ND, NB, E> dn = L::view(stages + ((j + NS - 2) % NS) * L::stage_bytes);
      const int t4 = threadIdx.x * C;
      op.row(dn, ce, up, t4, cons.p.x0 + t4, cons.yy - 1, t4 < cons.p.w);
      // the row two behind is done with; at the end of a tile so are the last two
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(empty + ((j + NS - 2) % NS));
        if (cons.yy == cons.p.y1) {
          mbar_arrive(empty + ((j + NS - 1) % NS));
          mbar_arrive(empty + (j % NS));
        }
      }
    }
    cons.next(g, T, th, list);
  }
}

#endif  // __CUDACC__

template <int ND, int NB, int NS, class T = double>
constexpr int smem_bytes() { return NS * Layout<ND, NB, T>::stage_bytes + 2 * NS * 8; }

}  // namespace pipe
}  // namespace euler
