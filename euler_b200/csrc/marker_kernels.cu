// euler_b200/csrc/marker_kernels.cu — marker particles: advection, re-binning / cell
// classification, deletion in sinks and solids, and fluid sources.
//
//   k_advect_markers   advect_markers + velocity_at + time_to   reference main.c:440-537
//   k_bin_markers      refresh_marker_counts (binning half)      main.c:102-117
//   k_seg_scan, k_list_deleted, k_fill_holes
//                      the swap-with-last deletion of main.c:112, done in parallel but
//                      producing the SAME array order as the sequential loop
//   k_fold_counts      uint32 atomic bins -> the reference's wrapping uint8 plane (main.c:96,114)
//   k_sources          update_fluid_sources + randf + xorshift64*  main.c:276-298, 203-207,
//                      misc/rng.c:5-20
//
// Positions are absolute fp32 world coordinates exactly as in the reference (no slab-local
// re-basing: SURVEY §7 hard part 8); all arithmetic is fp32 without FMA contraction in the
// reference's order, so positions — and therefore cell classification — are bit-exact.
#include "kernels.h"
#include "rng.cuh"
#include "marker_walk.cuh"

#include <float.h>

namespace euler {

namespace {

constexpr int SEG = 1024;            // markers per compaction segment
constexpr int MTHREADS = 256;

// ------------------------------------------------------------------ advection ----

__global__ void __launch_bounds__(MTHREADS) k_advect_markers(
    Grid g, InterpLimits lim, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid, float h,
    const float2* __restrict__ src, float2* __restrict__ dst, const DevScalars* sc, float dt) {
  const size_t n = sc->n_markers;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    dst[i] = walk_marker<false>(g, lim, u, v, fluid, solid, h, src[i], dt);
  }
}

// Row-slab mode: the same step, and a marker whose new cell row lies outside the rows this rank
// owns is handed to the neighbouring slab right here — appended to the staging buffer NCCL sends
// (displacement per sub-step is < 1 cell, CFL 0.75, main.c:838, so only to an ADJACENT slab; a
// handful of markers per boundary column, so one atomic each is cheap) and replaced in place by
// a position in the sink column x = 0 (main.c:244-252), where refresh_marker_counts' ordinary
// swap-delete removes it.  No separate partition pass over all markers.
__global__ void __launch_bounds__(MTHREADS) k_advect_markers_slab(
    Grid g, InterpLimits lim, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid, float h,
    const float2* __restrict__ src, float2* __restrict__ dst, DevScalars* sc, float dt,
    int own_lo, int own_hi, float2* __restrict__ send_dn, float2* __restrict__ send_up, size_t send_cap) {
  const size_t n = sc->n_markers;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float2 p = walk_marker<false>(g, lim, u, v, fluid, solid, h, src[i], dt);
    const int gy = (int)floorf(div_h(p.y, h));
    float2* out = gy < own_lo ? send_dn : (gy >= own_hi ? send_up : nullptr);   // null at the grid's ends too
    if (out) {
      const unsigned long long pos = atomicAdd(gy < own_lo ? &sc->n_send_dn : &sc->n_send_up, 1ull);
      if (pos < send_cap) out[pos] = p; else sc->marker_overflow = 1;
      p = make_float2(-1.f, -1.f);
    }
    dst[i] = p;
  }
}

__global__ void k_add_markers(DevScalars* sc, unsigned long long n) { sc->n_markers += n; }

// ---- reference marker mode: the `dt -= t_prev` carry-over (main.c:464, 501, 518) ---------
// In the reference `dt` is advect_markers' PARAMETER, so what one marker's rewind takes off
// it is also missing for every later marker of the array.  A walk with a smaller dt is a
// prefix of the walk with a larger one, so:
//   pass 1  every marker walks with the full dt0; markers whose rewind had t_prev > 0
//           ("candidates", rare: they crossed a cell edge and then hit a wall) are counted
//           per 1024-marker segment;
//   pass 2  ordered list of the candidates with their (t_hit, t_prev) records;
//   pass 3  ONE thread replays the reference's sequential bookkeeping over that short list:
//           which rewinds still fire under the shrinking dt, and dt after each candidate;
//   pass 4  every marker behind the first fired rewind re-walks with the dt it really had.
// All fp32 operations (dt - t_prev, t_hit < dt) happen in the reference's order, so the
// result is bit-identical to the sequential loop.
struct Candidate {
  unsigned int index;
  int n;
  float t_hit[2], t_prev[2];
};

__device__ __forceinline__ bool leaks(const HitRec& r) {
  return (r.n > 0 && r.t_prev[0] > 0.f) || (r.n > 1 && r.t_prev[1] > 0.f);
}

__global__ void __launch_bounds__(MTHREADS) k_advect_markers_ref(
    Grid g, InterpLimits lim, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid, float h,
    const float2* __restrict__ src, float2* __restrict__ dst, unsigned int* __restrict__ seg_count,
    DevScalars* sc, float dt) {
  const size_t n = sc->n_markers;
  const size_t nseg = (n + SEG - 1) / SEG;
  __shared__ int sh[MTHREADS / 32];
  for (size_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    int mine = 0;
#pragma unroll
    for (int k = 0; k < SEG / MTHREADS; ++k) {
      const size_t i = seg * SEG + (size_t)k * MTHREADS + threadIdx.x;
      if (i < n) {
        HitRec rec;
        dst[i] = walk_marker<true>(g, lim, u, v, fluid, solid, h, src[i], dt, &rec);
        mine += leaks(rec) ? 1 : 0;
      }
    }
    int w = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(EULER_FULL_MASK, w, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
#pragma unroll
      for (int k = 0; k < MTHREADS / 32; ++k) t += sh[k];
      seg_count[seg] = (unsigned int)t;
      if (t) atomicAdd(&sc->n_candidates, (unsigned long long)t);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(MTHREADS) k_list_candidates(
    Grid g, InterpLimits lim, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid, float h,
    const float2* __restrict__ src, const unsigned int* __restrict__ seg_count,
    const unsigned int* __restrict__ seg_offset, Candidate* __restrict__ cand, size_t cand_cap,
    const DevScalars* sc, float dt) {
  if (sc->n_candidates == 0 || sc->n_candidates > cand_cap) return;
  const size_t n = sc->n_markers;
  const size_t nseg = (n + SEG - 1) / SEG;
  __shared__ unsigned int warp_tot[MTHREADS / 32];
  for (size_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    if (seg_count[seg] == 0) continue;
    const unsigned int base = seg_offset[seg];
    HitRec rec[SEG / MTHREADS];
    bool is[SEG / MTHREADS];
    int mine = 0;
#pragma unroll
    for (int k = 0; k < SEG / MTHREADS; ++k) {
      const size_t i = seg * SEG + (size_t)threadIdx.x * (SEG / MTHREADS) + k;
      is[k] = false;
      if (i < n) {
        walk_marker<true>(g, lim, u, v, fluid, solid, h, src[i], dt, &rec[k]);
        is[k] = leaks(rec[k]);
      }
      mine += is[k];
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(EULER_FULL_MASK, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned int woff = 0;
    for (int wv = 0; wv < (threadIdx.x >> 5); ++wv) woff += warp_tot[wv];
    unsigned int rank = base + woff + incl - mine;
#pragma unroll
    for (int k = 0; k < SEG / MTHREADS; ++k) {
      if (is[k]) {
        Candidate c;
        c.index = (unsigned int)(seg * SEG + (size_t)threadIdx.x * (SEG / MTHREADS) + k);
        c.n = rec[k].n;
        c.t_hit[0] = rec[k].t_hit[0]; c.t_hit[1] = rec[k].t_hit[1];
        c.t_prev[0] = rec[k].t_prev[0]; c.t_prev[1] = rec[k].t_prev[1];
        cand[rank++] = c;
      }
    }
    __syncthreads();
  }
}

// The reference's sequential bookkeeping, replayed over the candidates only.
// dt_after[k] = value of `dt` after candidate k was processed; first_fired = index of the
// first candidate that actually changed dt (markers up to and including it keep dt0).
__global__ void __launch_bounds__(32) k_resolve_dt(const Candidate* __restrict__ cand,
                                                   float* __restrict__ dt_after, size_t cand_cap,
                                                   DevScalars* sc, float dt0) {
  // one warp: lanes fetch 32 records at a time (coalesced), then every lane replays them in
  // order from warp broadcasts, so the sequential chain never waits on global memory
  const unsigned long long nc = sc->n_candidates;
  const int lane = threadIdx.x;
  unsigned long long first = ~0ull;
  if (nc > cand_cap) { if (lane == 0) { sc->marker_overflow = 1; sc->first_fired = ~0ull; } return; }
  float run = dt0;
  for (unsigned long long base = 0; base < nc; base += 32) {
    Candidate mine;
    mine.index = 0; mine.n = 0;
    mine.t_hit[0] = mine.t_hit[1] = mine.t_prev[0] = mine.t_prev[1] = 0.f;
    if (base + lane < nc) mine = cand[base + lane];
    float my_after = 0.f;
    const int cnt = (int)((nc - base) < 32 ? (nc - base) : 32);
    for (int j = 0; j < cnt; ++j) {
      const int n = __shfl_sync(EULER_FULL_MASK, mine.n, j);
      const float h0 = __shfl_sync(EULER_FULL_MASK, mine.t_hit[0], j);
      const float h1 = __shfl_sync(EULER_FULL_MASK, mine.t_hit[1], j);
      const float p0 = __shfl_sync(EULER_FULL_MASK, mine.t_prev[0], j);
      const float p1 = __shfl_sync(EULER_FULL_MASK, mine.t_prev[1], j);
      const unsigned int idx = __shfl_sync(EULER_FULL_MASK, mine.index, j);
      float d = run;
      if (n > 0 && h0 < d) {                                  // `while (t_near < dt)` ... `dt -= t_prev`
        d -= p0;
        if (n > 1 && h1 < d) d -= p1;
      }
      if (d != run && first == ~0ull) first = idx;
      run = d;
      if (lane == j) my_after = run;
    }
    if (base + lane < nc) dt_after[base + lane] = my_after;
  }
  if (lane == 0) sc->first_fired = first;
}

__global__ void __launch_bounds__(MTHREADS) k_advect_fixup(
    Grid g, InterpLimits lim, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid, float h,
    const float2* __restrict__ src, float2* __restrict__ dst,
    const unsigned int* __restrict__ seg_count, const unsigned int* __restrict__ seg_offset,
    const Candidate* __restrict__ cand, const float* __restrict__ dt_after, DevScalars* sc,
    float dt0) {
  const unsigned long long first = sc->first_fired;
  if (sc->n_candidates == 0 || first == ~0ull) return;
  const size_t n = sc->n_markers;
  const size_t nseg = (n + SEG - 1) / SEG;
  for (size_t seg = first / SEG + blockIdx.x; seg < nseg; seg += gridDim.x) {
    const unsigned int off = seg_offset[seg], cnt = seg_count[seg];
    const float dt_start = off == 0 ? dt0 : dt_after[off - 1];
#pragma unroll
    for (int k = 0; k < SEG / MTHREADS; ++k) {
      const size_t i = seg * SEG + (size_t)k * MTHREADS + threadIdx.x;
      if (i >= n || i <= first) continue;
      float dt = dt_start;
      for (unsigned int j = 0; j < cnt && cand[off + j].index < i; ++j) dt = dt_after[off + j];
      if (dt != dt0) dst[i] = walk_marker<false>(g, lim, u, v, fluid, solid, h, src[i], dt);
    }
  }
}

// ------------------------------------------------------------------- binning ----

// One block per 1024-marker segment (grid-stride over segments).  Live markers are counted
// into the uint32 plane; markers in sink/solid cells are only counted per segment here.
__global__ void __launch_bounds__(MTHREADS) k_bin_markers(
    Grid g, float h, const float2* __restrict__ markers, const uint8_t* __restrict__ sink,
    const uint8_t* __restrict__ solid, unsigned int* __restrict__ count32,
    unsigned int* __restrict__ seg_count, DevScalars* sc) {
  const size_t n = sc->n_markers;
  const size_t nseg = (n + SEG - 1) / SEG;
  for (size_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    int dead = 0;
#pragma unroll
    for (int k = 0; k < SEG / MTHREADS; ++k) {
      const size_t i = seg * SEG + (size_t)k * MTHREADS + threadIdx.x;
      // (electing one lane per cell with __match_any_sync — markers of a cell are mostly adjacent —
      // was measured: refresh_counts 1.00 -> 2.01 ms at 16384^2; the L2 handles the same-address adds
      // of a warp faster than the match does)
      if (i < n) {
        size_t c;
        marker_cell(g, h, markers[i], &c);
        if (sink[c] || solid[c]) ++dead;
        else atomicAdd(count32 + c, 1u);
      }
    }
    // block-wide sum of `dead`
    __shared__ int sh[MTHREADS / 32];
    int w = dead;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(EULER_FULL_MASK, w, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
#pragma unroll
      for (int k = 0; k < MTHREADS / 32; ++k) s += sh[k];
      seg_count[seg] = (unsigned int)s;
      if (s) atomicAdd(&sc->n_deleted, (unsigned long long)s);
    }
    __syncthreads();
  }
}

// Exclusive scan of seg_count (single block, 1024 threads, contiguous chunk per thread).
__global__ void __launch_bounds__(1024) k_seg_scan(const unsigned int* __restrict__ seg_count,
                                                   unsigned int* __restrict__ seg_offset,
                                                   const DevScalars* sc, int for_candidates) {
  if ((for_candidates ? sc->n_candidates : sc->n_deleted) == 0) return;
  const size_t n = sc->n_markers;
  const size_t nseg = (n + SEG - 1) / SEG;
  const size_t per = (nseg + 1023) / 1024;
  const size_t lo = min(nseg, per * threadIdx.x), hi = min(nseg, lo + per);
  unsigned int sum = 0;
  for (size_t i = lo; i < hi; ++i) sum += seg_count[i];
  __shared__ unsigned int sh[1024];
  sh[threadIdx.x] = sum;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 thread sums
  for (int d = 1; d < 1024; d <<= 1) {
    unsigned int vv = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
    __syncthreads();
    sh[threadIdx.x] += vv;
    __syncthreads();
  }
  unsigned int run = sh[threadIdx.x] - sum;
  for (size_t i = lo; i < hi; ++i) { seg_offset[i] = run; run += seg_count[i]; }
}

// Ascending list of the indices of deleted markers.
__global__ void __launch_bounds__(MTHREADS) k_list_deleted(
    Grid g, float h, const float2* __restrict__ markers, const uint8_t* __restrict__ sink,
    const uint8_t* __restrict__ solid, const unsigned int* __restrict__ seg_count,
    const unsigned int* __restrict__ seg_offset, unsigned int* __restrict__ del_list,
    const DevScalars* sc) {
  if (sc->n_deleted == 0) return;
  const size_t n = sc->n_markers;
  const size_t nseg = (n + SEG - 1) / SEG;
  __shared__ unsigned int warp_tot[MTHREADS / 32];
  for (size_t seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    if (seg_count[seg] == 0) continue;                       // block-uniform
    unsigned int base = seg_offset[seg];
    // index order inside a segment: thread t owns markers seg*SEG + 4t .. 4t+3
    bool dead[SEG / MTHREADS];
    int mine = 0;
#pragma unroll
    for (int k = 0; k < SEG / MTHREADS; ++k) {
      const size_t i = seg * SEG + (size_t)threadIdx.x * (SEG / MTHREADS) + k;
      dead[k] = false;
      if (i < n) {
        size_t c;
        marker_cell(g, h, markers[i], &c);
        dead[k] = sink[c] || solid[c];
      }
      mine += dead[k];
    }
    // exclusive scan of `mine` across the block
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(EULER_FULL_MASK, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned int woff = 0;
    for (int wv = 0; wv < (threadIdx.x >> 5); ++wv) woff += warp_tot[wv];
    unsigned int rank = base + woff + incl - mine;
#pragma unroll
    for (int k = 0; k < SEG / MTHREADS; ++k) {
      if (dead[k]) {
        del_list[rank++] = (unsigned int)(seg * SEG + (size_t)threadIdx.x * (SEG / MTHREADS) + k);
      }
    }
    __syncthreads();
  }
}

// The sequential loop `if dead: m[i--] = m[--len]` (main.c:105-116) leaves survivors below
// the final length S in place and fills the k-th hole (ascending) with the k-th survivor
// counted from the END of the array (descending).  Single block; the tail [S, n) is as long
// as the number of deletions.
__global__ void __launch_bounds__(1024) k_fill_holes(
    Grid g, float h, float2* markers, const uint8_t* __restrict__ sink,
    const uint8_t* __restrict__ solid, const unsigned int* __restrict__ del_list,
    DevScalars* sc) {
  const unsigned long long nd = sc->n_deleted;
  if (nd == 0) return;
  const size_t n = sc->n_markers;
  const size_t S = n - nd;
  __shared__ unsigned int warp_tot[32];
  __shared__ unsigned int carry_sh;
  if (threadIdx.x == 0) carry_sh = 0;
  __syncthreads();
  // walk the tail from the end in chunks of 1024: position j = n-1 - (chunk*1024 + tid)
  for (size_t done = 0; done < nd; done += 1024) {
    const size_t back = done + threadIdx.x;
    bool live = false;
    float2 m = make_float2(0.f, 0.f);
    if (back < nd) {
      const size_t j = n - 1 - back;
      m = markers[j];
      size_t c;
      marker_cell(g, h, m, &c);
      live = !(sink[c] || solid[c]);
    }
    const unsigned int bal = __ballot_sync(EULER_FULL_MASK, live);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    unsigned int woff = 0, total = 0;
    for (int wv = 0; wv < 32; ++wv) {
      const unsigned int t = warp_tot[wv];
      if (wv < wid) woff += t;
      total += t;
    }
    const unsigned int rank = carry_sh + woff + __popc(bal & ((1u << lane) - 1u));
    if (live) markers[del_list[rank]] = m;                  // hole index < S <= every tail index
    __syncthreads();
    if (threadIdx.x == 0) carry_sh += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sc->n_markers = S;
    sc->n_deleted = 0;
  }
}

// Over the tiles of the LAST sub-step's list (common.cuh GridTiles; all tiles when there is none):
// a marker moves less than one cell per sub-step (CFL 0.75, main.c:838), so every cell that was
// binned into, and every stale cell of the count plane being recycled (the one of two sub-steps
// ago), lies in a tile that held or bordered fluid within the list's three-sub-step memory.
__global__ void __launch_bounds__(256) k_fold_counts(Grid g, GridTiles gt, unsigned int* __restrict__ count32,
                                                     uint8_t* __restrict__ count) {
  const unsigned int n = gt.list ? *gt.count : (unsigned int)(gt.tx * gt.ty);
  constexpr int QPT = GT_W / 4;                               // quads per tile row
  for (unsigned int t = blockIdx.x; t < n; t += gridDim.x) {
    const int tile = gt.list ? gt.list[t] : (int)t;
    const int x00 = (tile % gt.tx) * GT_W, y0 = (tile / gt.tx) * GT_H;
    for (int i = threadIdx.x; i < QPT * GT_H; i += blockDim.x) {
      const int y = y0 + i / QPT, x0 = x00 + (i % QPT) * 4;
      if (y >= g.ny || x0 >= g.pitch) continue;
      const size_t c = gidx(g, x0, y);
      uint4 v = *reinterpret_cast<uint4*>(count32 + c);
      uchar4 o = make_uchar4((unsigned char)v.x, (unsigned char)v.y, (unsigned char)v.z,
                             (unsigned char)v.w);            // uint8 wrap, main.c:96,114
      *reinterpret_cast<uchar4*>(count + c) = o;
      if (v.x | v.y | v.z | v.w) *reinterpret_cast<uint4*>(count32 + c) = make_uint4(0, 0, 0, 0);
    }
  }
}


// ---- row-slab mode: markers that left the rows this rank owns -------------------------------
// Displacement per sub-step is < 1 cell (CFL 0.75, main.c:838), so a marker can only move to
// an adjacent slab.  Out-of-place three-way partition with warp-aggregated cursors: keepers
// go to `keep`, leavers to the staging buffers that NCCL sends to the neighbours.
__global__ void __launch_bounds__(MTHREADS) k_partition_markers(
    float h, int own_lo, int own_hi, const float2* __restrict__ src, float2* __restrict__ keep,
    float2* __restrict__ send_dn, float2* __restrict__ send_up, size_t send_cap, DevScalars* sc,
    unsigned long long* n_keep) {
  const size_t n = sc->n_markers;
  const size_t per = (size_t)gridDim.x * blockDim.x;
  const size_t rounds = (n + per - 1) / per;
  const int lane = threadIdx.x & 31;
  for (size_t rd = 0; rd < rounds; ++rd) {
    const size_t i = rd * per + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    int cls = -1;                                     // 0 keep, 1 down, 2 up
    float2 m = make_float2(0.f, 0.f);
    if (i < n) {
      m = src[i];
      const int gy = (int)floorf(div_h(m.y, h));
      cls = gy < own_lo ? 1 : (gy >= own_hi ? 2 : 0);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const unsigned bal = __ballot_sync(EULER_FULL_MASK, cls == k);
      if (!bal) continue;
      unsigned long long base = 0;
      if (lane == __ffs(bal) - 1)
        base = atomicAdd(k == 0 ? n_keep : (k == 1 ? &sc->n_send_dn : &sc->n_send_up),
                         (unsigned long long)__popc(bal));
      base = __shfl_sync(EULER_FULL_MASK, base, __ffs(bal) - 1);
      if (cls == k) {
        const unsigned long long pos = base + __popc(bal & ((1u << lane) - 1u));
        if (k == 0) keep[pos] = m;
        else if (pos < send_cap) (k == 1 ? send_dn : send_up)[pos] = m;
        else sc->marker_overflow = 1;
      }
    }
  }
}

// create()/reinit() of a slab handle: of `n` staged markers, append those whose cell row is
// in [own_lo, own_hi) to `dst` at sc->n_markers (warp-aggregated append; order is free in
// FAST marker mode).  Markers beyond `cap` are counted but not stored: the caller checks.
__global__ void __launch_bounds__(MTHREADS) k_filter_markers(
    float h, int own_lo, int own_hi, const float2* __restrict__ src, size_t n,
    float2* __restrict__ dst, size_t cap, DevScalars* sc) {
  // one atomic per BLOCK and round (a warp-level cursor on one address serialises: 3 ms for 10^8
  // markers): warp ballots, block scan of the warp counts, thread 0 reserves the block's range
  const size_t per = (size_t)gridDim.x * blockDim.x;
  const size_t rounds = (n + per - 1) / per;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __shared__ unsigned int wcount[MTHREADS / 32];
  __shared__ unsigned long long base_sh;
  for (size_t rd = 0; rd < rounds; ++rd) {
    const size_t i = rd * per + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    bool mine = false;
    float2 m = make_float2(0.f, 0.f);
    if (i < n) {
      m = src[i];
      const int gy = (int)floorf(div_h(m.y, h));
      mine = gy >= own_lo && gy < own_hi;
    }
    const unsigned bal = __ballot_sync(EULER_FULL_MASK, mine);
    if (lane == 0) wcount[wid] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int tot = 0;
#pragma unroll
      for (int k = 0; k < MTHREADS / 32; ++k) { const unsigned int t = wcount[k]; wcount[k] = tot; tot += t; }
      base_sh = tot ? atomicAdd(&sc->n_markers, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    if (mine) {
      const unsigned long long pos = base_sh + wcount[wid] + __popc(bal & ((1u << lane) - 1u));
      if (pos < cap) dst[pos] = m;
    }
    __syncthreads();
  }
}

// source cells of this slab that need a marker (main.c:287), for the cross-rank prefix
__global__ void __launch_bounds__(1024) k_sources_count(const unsigned int* __restrict__ cells,
                                                        size_t ncells, const uint8_t* __restrict__ count,
                                                        DevScalars* sc) {
  int mine = 0;
  for (size_t i = threadIdx.x; i < ncells; i += 1024) mine += count[cells[i]] < 4 ? 1 : 0;
  __shared__ int sh[32];
  int w = mine;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(EULER_FULL_MASK, w, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < 32; ++k) t += sh[k];
    sc->part[2] = (double)t;
    sc->part[3] = (double)sc->n_markers;
  }
}

// fold the all-gathered {need, n_markers} pairs: my base rank in the global row-major order of
// needy source cells, the global marker count and the global number of needy cells
__global__ void k_sources_prep(DevScalars* sc, const double* __restrict__ gathered, int rank, int nranks) {
  double base = 0.0, total = 0.0, nglob = 0.0;
  for (int r = 0; r < nranks; ++r) {
    if (r < rank) base += gathered[r * 4 + 2];
    total += gathered[r * 4 + 2];
    nglob += gathered[r * 4 + 3];
  }
  sc->src_base = (unsigned long long)base;
  sc->n_markers_global = (unsigned long long)nglob;
  sc->part[2] = total;
}

// ------------------------------------------------------------------- sources ----

// Row-major over source cells; the k-th cell that needs a marker uses draws 2k and 2k+1 of
// the stream (y jitter first: gcc evaluates v2f's second argument first, see oracle).
// Three passes so that a scenario with 10^5..10^6 source cells (waterfall at 4096^2: 270 000) does
// not run on one block (it took 1.25 ms of a 6 ms sub-step there): (1) needy cells per block of
// 1024 source cells, (2) one block scans those counts and does the sequential loop's bookkeeping
// (marker total, RNG state after all draws, the MAX_MARKER_COUNT-1 latch), (3) every block emits
// its markers at rank = cells-before-me, each with its own jump-ahead of the stream.
constexpr int SRC_CHUNK = 1024;

__global__ void __launch_bounds__(SRC_CHUNK) k_sources_need(const unsigned int* __restrict__ cells, size_t ncells,
                                                            const uint8_t* __restrict__ count,
                                                            unsigned int* __restrict__ block_need) {
  const size_t i = (size_t)blockIdx.x * SRC_CHUNK + threadIdx.x;
  const bool need = i < ncells && count[cells[i]] < 4;       // main.c:287
  const int n = __syncthreads_count(need);
  if (threadIdx.x == 0) block_need[blockIdx.x] = (unsigned int)n;
}

__global__ void __launch_bounds__(1024) k_sources_scan(const unsigned int* __restrict__ block_need,
                                                       unsigned int* __restrict__ block_base, int nblocks,
                                                       size_t max_markers_global,
                                                       const unsigned long long* __restrict__ jump, DevScalars* sc,
                                                       int distributed) {
  // exclusive scan of the per-block counts (contiguous chunk per thread + Hillis-Steele over 1024 sums)
  const int per = (nblocks + 1023) / 1024;
  const int lo = min(nblocks, per * (int)threadIdx.x), hi = min(nblocks, lo + per);
  unsigned int sum = 0;
  for (int i = lo; i < hi; ++i) sum += block_need[i];
  __shared__ unsigned int sh[1024];
  sh[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const unsigned int v = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
    __syncthreads();
    sh[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned int run = sh[threadIdx.x] - sum;
  for (int i = lo; i < hi; ++i) { block_base[i] = run; run += block_need[i]; }
  if (threadIdx.x != 1023) return;
  // single GPU: ranks start at 0 and the global marker count is the local one.  Slab mode:
  // this slab's needy cells come after those of the slabs below it (row-major order), the
  // MAX_MARKER_COUNT-1 latch (main.c:281,290) looks at the global count.
  const unsigned long long carry = sh[1023];
  const unsigned long long n0 = sc->n_markers;
  const unsigned long long nglob = distributed ? sc->n_markers_global : n0;
  const unsigned long long rank0 = distributed ? sc->src_base : 0ull;
  const unsigned long long cap = max_markers_global - 1;     // main.c:281
  const bool exhausted0 = sc->source_exhausted || nglob == cap;
  const unsigned long long allow = exhausted0 ? 0 : cap - nglob;
  const unsigned long long state0 = sc->rng_state;
  sc->src_n0 = n0; sc->src_state0 = state0; sc->src_allow = allow;
  // markers this slab appends: its needy cells whose global rank is below `allow`
  unsigned long long mine = carry;
  if (rank0 >= allow) mine = 0;
  else if (rank0 + mine > allow) mine = allow - rank0;
  const unsigned long long need_all = distributed ? (unsigned long long)sc->part[2] : carry;
  const unsigned long long added_all = need_all < allow ? need_all : allow;
  sc->n_markers = n0 + mine;
  sc->rng_state = rng_jump(jump, state0, 2 * added_all);
  sc->source_exhausted = (exhausted0 || (nglob + added_all == cap)) ? 1 : 0;
}

__global__ void __launch_bounds__(SRC_CHUNK) k_sources_emit(
    Grid g, float h, const unsigned int* __restrict__ cells, size_t ncells, uint8_t* __restrict__ count,
    float2* __restrict__ markers, const unsigned int* __restrict__ block_base,
    const unsigned long long* __restrict__ jump, const DevScalars* sc, int distributed) {
  const unsigned long long n0 = sc->src_n0, state0 = sc->src_state0, allow = sc->src_allow;
  const unsigned long long rank0 = distributed ? sc->src_base : 0ull;
  const size_t i = (size_t)blockIdx.x * SRC_CHUNK + threadIdx.x;
  bool need = false;
  unsigned int cell = 0;
  if (i < ncells) {
    cell = cells[i];
    need = count[cell] < 4;                                  // main.c:287
  }
  __shared__ unsigned int warp_tot[SRC_CHUNK / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned int bal = __ballot_sync(EULER_FULL_MASK, need);
  if (lane == 0) warp_tot[wid] = __popc(bal);
  __syncthreads();
  unsigned int woff = 0;
  for (int wv = 0; wv < wid; ++wv) woff += warp_tot[wv];
  const unsigned long long lrank = (unsigned long long)block_base[blockIdx.x] + woff + __popc(bal & ((1u << lane) - 1u));
  const unsigned long long rank = rank0 + lrank;
  if (need && rank < allow) {
    unsigned long long s = rng_jump(jump, state0, 2 * rank);
    s = rng_step(s);
    const float jy = rng_float(s);
    s = rng_step(s);
    const float jx = rng_float(s);
    const int y = (int)(cell / (unsigned int)g.pitch), x = (int)(cell % (unsigned int)g.pitch);
    markers[n0 + lrank] = make_float2(h * (x + jx), h * (y + g.yoff + jy));   // main.c:288
    count[cell] += 1;
  }
}

}  // namespace

void init_rng_jump_table(unsigned long long* t) { rng_build_jump_table(t); }

void launch_advect_markers(Ctx& c, float dt, int mode) {
  ProfScope ps(c, KC_ADVECT_MARKERS);
  const int blocks = c.sm_count * 8;
  if (mode == 1) {          // EULER_MARKERS_FAST: every marker gets the full dt
    k_advect_markers<<<blocks, MTHREADS, 0, c.stream>>>(c.g, c.lim, c.u, c.v, c.count, c.solid,
                                                        c.h, c.markers, c.markers_alt, c.sc, dt);
    c.launches += 1;
  } else {                  // EULER_MARKERS_REFERENCE: reproduce the dt carry-over
    Candidate* cand = reinterpret_cast<Candidate*>(c.cand);
    cudaMemsetAsync(&c.sc->n_candidates, 0, sizeof(unsigned long long), c.stream);
    k_advect_markers_ref<<<blocks, MTHREADS, 0, c.stream>>>(
        c.g, c.lim, c.u, c.v, c.count, c.solid, c.h, c.markers, c.markers_alt, c.seg_count, c.sc, dt);
    k_seg_scan<<<1, 1024, 0, c.stream>>>(c.seg_count, c.seg_offset, c.sc, 1);
    k_list_candidates<<<blocks, MTHREADS, 0, c.stream>>>(
        c.g, c.lim, c.u, c.v, c.count, c.solid, c.h, c.markers, c.seg_count, c.seg_offset, cand,
        c.cand_cap, c.sc, dt);
    k_resolve_dt<<<1, 32, 0, c.stream>>>(cand, c.cand_dt, c.cand_cap, c.sc, dt);
    k_advect_fixup<<<blocks, MTHREADS, 0, c.stream>>>(
        c.g, c.lim, c.u, c.v, c.count, c.solid, c.h, c.markers, c.markers_alt, c.seg_count,
        c.seg_offset, cand, c.cand_dt, c.sc, dt);
    c.launches += 5;
  }
  float2* t = c.markers; c.markers = c.markers_alt; c.markers_alt = t;
}

size_t marker_candidate_bytes() { return sizeof(Candidate); }

void launch_refresh_counts(Ctx& c) {
  ProfScope ps(c, KC_REFRESH_COUNTS);
  // prev <- cur (main.c:103) by swapping planes; cur is rebuilt from scratch (main.c:104)
  uint8_t* t = c.prev_count; c.prev_count = c.count; c.count = t;
  const int blocks = c.sm_count * 8;
  k_bin_markers<<<blocks, MTHREADS, 0, c.stream>>>(c.g, c.h, c.markers, c.sink, c.solid,
                                                   c.count32, c.seg_count, c.sc);
  k_seg_scan<<<1, 1024, 0, c.stream>>>(c.seg_count, c.seg_offset, c.sc, 0);
  unsigned int* del_list = reinterpret_cast<unsigned int*>(c.markers_alt);
  k_list_deleted<<<blocks, MTHREADS, 0, c.stream>>>(c.g, c.h, c.markers, c.sink, c.solid,
                                                    c.seg_count, c.seg_offset, del_list, c.sc);
  k_fill_holes<<<1, 1024, 0, c.stream>>>(c.g, c.h, c.markers, c.sink, c.solid, del_list, c.sc);
  k_fold_counts<<<blocks, 256, 0, c.stream>>>(c.g, grid_tiles_of(c, c.gt_prev_sparse), c.count32, c.count);
  c.launches += 5;
}

// ---- the static row-major list of source cells (main.c:284-286 visits them in this order), built
// on the device from the uploaded source plane: per-row counts, one scan over the rows, ordered
// writes.  (Round 1 scanned the plane on the host: 35 ms per sim_init hand-over at 16384^2.)
namespace {
__device__ __forceinline__ int nonzero_bytes(unsigned w) {
  return __popc((((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u);
}
// one block per stored row; padding columns of the plane are zero
__global__ void __launch_bounds__(256) k_src_row_count(Grid g, const uint8_t* __restrict__ source,
                                                       unsigned int* __restrict__ row_count) {
  __shared__ int tot;
  if (threadIdx.x == 0) tot = 0;
  __syncthreads();
  const uint4* row = reinterpret_cast<const uint4*>(source + (size_t)blockIdx.x * g.pitch);
  int n = 0;
  for (int i = threadIdx.x; i < g.pitch / 16; i += blockDim.x) {
    const uint4 w = row[i];
    if (w.x | w.y | w.z | w.w) n += nonzero_bytes(w.x) + nonzero_bytes(w.y) + nonzero_bytes(w.z) + nonzero_bytes(w.w);
  }
  if (n) atomicAdd(&tot, n);
  __syncthreads();
  if (threadIdx.x == 0) row_count[blockIdx.x] = (unsigned int)tot;
}
// exclusive scan of the OWNED rows' counts (in place), totals[0] = cells in owned rows,
// totals[1] = cells in all stored rows
__global__ void __launch_bounds__(1024) k_src_row_scan(int ny, int own0, int own1, unsigned int* __restrict__ row_count,
                                                       unsigned long long* __restrict__ totals) {
  __shared__ unsigned long long sh[1024];
  __shared__ unsigned long long all_sh;
  const int per = (ny + 1023) / 1024;
  const int lo = min(ny, per * (int)threadIdx.x), hi = min(ny, lo + per);
  unsigned long long own = 0, all = 0;
  for (int y = lo; y < hi; ++y) { all += row_count[y]; if (y >= own0 && y < own1) own += row_count[y]; }
  if (threadIdx.x == 0) all_sh = 0;
  sh[threadIdx.x] = own;
  __syncthreads();
  if (all) atomicAdd(&all_sh, all);
  for (int d = 1; d < 1024; d <<= 1) {
    const unsigned long long v = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
    __syncthreads();
    sh[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned long long run = sh[threadIdx.x] - own;
  for (int y = lo; y < hi; ++y) {
    const unsigned int c = (y >= own0 && y < own1) ? row_count[y] : 0u;
    row_count[y] = c ? (unsigned int)run : 0xffffffffu;       // rows without (owned) sources are skipped
    run += c;
  }
  if (threadIdx.x == 1023) { totals[0] = sh[1023]; totals[1] = all_sh; }
}
// one block per stored row that has sources: cells in ascending x
__global__ void __launch_bounds__(256) k_src_row_write(Grid g, const uint8_t* __restrict__ source,
                                                       const unsigned int* __restrict__ row_offset,
                                                       unsigned int* __restrict__ cells) {
  const unsigned int base0 = row_offset[blockIdx.x];
  if (base0 == 0xffffffffu) return;
  __shared__ int wsum[8];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const unsigned* row = reinterpret_cast<const unsigned*>(source + (size_t)blockIdx.x * g.pitch);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i0 = 0; i0 < g.pitch / 4; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const unsigned w = i < g.pitch / 4 ? row[i] : 0u;
    const int n = nonzero_bytes(w);
    int inc = n;                                               // inclusive scan over the block
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(EULER_FULL_MASK, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int before = carry;
    for (int k = 0; k < wid; ++k) before += wsum[k];
    int pos = before + inc - n;
    for (int b = 0; b < 4; ++b)
      if ((w >> (8 * b)) & 0xffu) cells[base0 + pos++] = (unsigned int)((size_t)blockIdx.x * g.pitch + 4 * i + b);
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = before + inc;
    __syncthreads();
  }
}
}  // namespace

void launch_source_rows_count(Ctx& c, unsigned int* rows_scratch, unsigned long long* totals2) {
  k_src_row_count<<<c.g.ny, 256, 0, c.stream>>>(c.g, c.source, rows_scratch);
  k_src_row_scan<<<1, 1024, 0, c.stream>>>(c.g.ny, c.own0, c.own1, rows_scratch, totals2);
  c.launches += 2;
}
void launch_source_rows_write(Ctx& c, const unsigned int* rows_scratch) {
  k_src_row_write<<<c.g.ny, 256, 0, c.stream>>>(c.g, c.source, rows_scratch, c.source_cells);
  c.launches += 1;
}

void launch_sources(Ctx& c) {
  if (c.n_source_cells_global == 0) return;
  ProfScope ps(c, KC_SOURCES);
  // (seg_count / seg_offset: the compaction scratch of refresh_marker_counts, free at this point;
  // one entry per 1024 markers of capacity >= one per 1024 source cells)
  const int nb = (int)((c.n_source_cells + SRC_CHUNK - 1) / SRC_CHUNK);
  if (nb > 0) {
    k_sources_need<<<nb, SRC_CHUNK, 0, c.stream>>>(c.source_cells, c.n_source_cells, c.count, c.seg_count);
    c.launches += 1;
  }
  k_sources_scan<<<1, 1024, 0, c.stream>>>(c.seg_count, c.seg_offset, nb, c.max_markers_global, c.rng_jump, c.sc,
                                           c.distributed);
  c.launches += 1;
  if (nb > 0) {
    k_sources_emit<<<nb, SRC_CHUNK, 0, c.stream>>>(c.g, c.h, c.source_cells, c.n_source_cells, c.count, c.markers,
                                                   c.seg_offset, c.rng_jump, c.sc, c.distributed);
    c.launches += 1;
  }
}

void launch_sources_count(Ctx& c) {
  k_sources_count<<<1, 1024, 0, c.stream>>>(c.source_cells, c.n_source_cells, c.count, c.sc);
  c.launches += 1;
}
void launch_sources_prep(Ctx& c, const double* gathered, int rank, int nranks) {
  k_sources_prep<<<1, 1, 0, c.stream>>>(c.sc, gathered, rank, nranks);
  c.launches += 1;
}

// slab mode: advection + hand-over of the leavers in one pass (k_advect_markers_slab)
void launch_advect_markers_slab(Ctx& c, float dt, int own_lo_global, int own_hi_global, float2* send_dn,
                                float2* send_up, size_t send_cap) {
  ProfScope ps(c, KC_ADVECT_MARKERS);
  cudaMemsetAsync(&c.sc->n_send_dn, 0, 2 * sizeof(unsigned long long), c.stream);
  k_advect_markers_slab<<<c.sm_count * 8, MTHREADS, 0, c.stream>>>(
      c.g, c.lim, c.u, c.v, c.count, c.solid, c.h, c.markers, c.markers_alt, c.sc, dt, own_lo_global,
      own_hi_global, send_dn, send_up, send_cap);
  c.launches += 1;
  float2* t = c.markers; c.markers = c.markers_alt; c.markers_alt = t;
}
void launch_add_markers(Ctx& c, unsigned long long n) {
  k_add_markers<<<1, 1, 0, c.stream>>>(c.sc, n);
  c.launches += 1;
}

// three-way partition; keepers land in markers_alt, which becomes the marker array
void launch_partition_markers(Ctx& c, int own_lo_global, int own_hi_global, float2* send_dn,
                              float2* send_up, size_t send_cap, unsigned long long* n_keep) {
  ProfScope ps(c, KC_REFRESH_COUNTS);
  cudaMemsetAsync(&c.sc->n_send_dn, 0, 2 * sizeof(unsigned long long), c.stream);
  cudaMemsetAsync(n_keep, 0, sizeof(unsigned long long), c.stream);
  k_partition_markers<<<c.sm_count * 8, MTHREADS, 0, c.stream>>>(
      c.h, own_lo_global, own_hi_global, c.markers, c.markers_alt, send_dn, send_up, send_cap, c.sc, n_keep);
  c.launches += 1;
  float2* t = c.markers; c.markers = c.markers_alt; c.markers_alt = t;
}

void launch_filter_markers(Ctx& c, const float2* staged, size_t n, int own_lo_global, int own_hi_global) {
  if (!n) return;
  k_filter_markers<<<c.sm_count * 8, MTHREADS, 0, c.stream>>>(c.h, own_lo_global, own_hi_global, staged, n,
                                                              c.markers, c.max_markers, c.sc);
  c.launches += 1;
}

}  // namespace euler
