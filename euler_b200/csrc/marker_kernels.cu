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

#include <float.h>

namespace euler {

namespace {

constexpr int SEG = 1024;            // markers per compaction segment
constexpr int MTHREADS = 256;

// ------------------------------------------------------------------ advection ----

__device__ __forceinline__ float time_until(float from, float to, float vel) {
  return fabsf(vel) > 0.f ? (to - from) / vel : FLT_MAX;     // main.c:451-457
}

// One marker, RK1 with the reference's grid-line walk (main.c:466-535).
__device__ __forceinline__ float2 walk_marker(const Grid& g, const InterpLimits& lim,
                                              const float* __restrict__ u,
                                              const float* __restrict__ v,
                                              const uint8_t* __restrict__ fluid,
                                              const uint8_t* __restrict__ solid, float h,
                                              float2 pos, float dt) {
  float px = pos.x, py = pos.y;
  // velocity_at, main.c:440-449
  float vx = interpolate<FACE_U>(u, fluid, g, lim, px / h - 1.f, py / h - 0.5f);
  float vy = interpolate<FACE_V>(v, fluid, g, lim, px / h - 0.5f, py / h - 1.f);

  int cx = (int)floorf(px / h);
  int cy = (int)floorf(py / h);
  const int step_x = vx > 0 ? 1 : -1;
  const int step_y = vy > 0 ? 1 : -1;
  int line_x = cx + (vx > 0 ? 1 : 0);
  int line_y = cy + (vy > 0 ? 1 : 0);
  const int off_x = vx < 0 ? -1 : 0;
  const int off_y = vy < 0 ? -1 : 0;
  float gx = line_x * h, gy = line_y * h;
  float tx = time_until(px, gx, vx);
  float ty = time_until(py, gy, vy);

  float t_prev = 0.f;
  float t_next = fminf(tx, ty);
  while (t_next < dt) {
    if (tx < ty) {
      const int sx = min(max(line_x + off_x, 0), g.nx - 1), sy = min(max(cy, 0), g.ny - 1);
      if (solid[gidx(g, sx, sy)]) {
        px = px + t_prev * vx; py = py + t_prev * vy;        // rewind, main.c:500
        dt -= t_prev;
        t_next = 0.f;
        vx = 0.f;
        tx = FLT_MAX;
        ty = time_until(py, gy, vy);
      } else {
        cx = line_x;
        line_x = cx + step_x;
        gx = line_x * h;
        tx = time_until(px, gx, vx);
      }
    } else {
      const int sx = min(max(cx, 0), g.nx - 1), sy = min(max(line_y + off_y, 0), g.ny - 1);
      if (solid[gidx(g, sx, sy)]) {
        px = px + t_prev * vx; py = py + t_prev * vy;        // main.c:517
        dt -= t_prev;
        t_next = 0.f;
        vy = 0.f;
        ty = FLT_MAX;
        tx = time_until(px, gx, vx);
      } else {
        cy = line_y;
        line_y = cy + step_y;
        gy = line_y * h;
        ty = time_until(py, gy, vy);
      }
    }
    t_prev = t_next;
    t_next = fminf(tx, ty);
  }
  const float t = (t_next < FLT_MAX) ? dt : t_prev;          // main.c:534
  return make_float2(px + t * vx, py + t * vy);
}

__global__ void __launch_bounds__(MTHREADS) k_advect_markers(
    Grid g, InterpLimits lim, const float* __restrict__ u, const float* __restrict__ v,
    const uint8_t* __restrict__ fluid, const uint8_t* __restrict__ solid, float h,
    const float2* __restrict__ src, float2* __restrict__ dst, const DevScalars* sc, float dt) {
  const size_t n = sc->n_markers;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    dst[i] = walk_marker(g, lim, u, v, fluid, solid, h, src[i], dt);
  }
}

// ------------------------------------------------------------------- binning ----

__device__ __forceinline__ bool marker_cell(const Grid& g, float h, float2 p, size_t* cell) {
  int cx = (int)floorf(p.x / h);                             // main.c:106-107
  int cy = (int)floorf(p.y / h);
  // the reference asserts 0 < x < X, 0 < y < Y (main.c:108, compiled out); clamp so a stray
  // marker lands in the sink ring and is deleted instead of indexing out of bounds
  cx = min(max(cx, 0), g.nx - 1);
  cy = min(max(cy, 0), g.ny - 1);
  *cell = gidx(g, cx, cy);
  return true;
}

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
                                                   const DevScalars* sc) {
  if (sc->n_deleted == 0) return;
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

__global__ void __launch_bounds__(256) k_fold_counts(Grid g, unsigned int* __restrict__ count32,
                                                     uint8_t* __restrict__ count) {
  const int quads = g.pitch >> 2;
  const size_t total = (size_t)quads * g.ny;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / quads);
    const int x0 = (int)(i % quads) << 2;
    const size_t c = gidx(g, x0, y);
    uint4 v = *reinterpret_cast<uint4*>(count32 + c);
    uchar4 o = make_uchar4((unsigned char)v.x, (unsigned char)v.y, (unsigned char)v.z,
                           (unsigned char)v.w);            // uint8 wrap, main.c:96,114
    *reinterpret_cast<uchar4*>(count + c) = o;
    *reinterpret_cast<uint4*>(count32 + c) = make_uint4(0, 0, 0, 0);
  }
}

// ------------------------------------------------------------------- sources ----

// xorshift64 state transition is linear over GF(2) (misc/rng.c:7-9); jump[j] holds the 64
// columns of T^(2^j), so any number of draws can be skipped in O(64 log k).
__device__ __forceinline__ unsigned long long rng_jump(const unsigned long long* __restrict__ jump,
                                                       unsigned long long s, unsigned long long k) {
  for (int j = 0; k != 0; ++j, k >>= 1) {
    if (k & 1ull) {
      const unsigned long long* col = jump + (size_t)j * 64;
      unsigned long long out = 0;
      for (int b = 0; b < 64; ++b)
        if ((s >> b) & 1ull) out ^= col[b];
      s = out;
    }
  }
  return s;
}
__device__ __forceinline__ unsigned long long rng_step(unsigned long long s) {
  s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
  return s;
}
__device__ __forceinline__ float rng_float(unsigned long long s) {
  const unsigned int bits = (unsigned int)((s * 0x2545F4914F6CDD1Dull) >> 32);
  return (float)((double)bits / (double)0xFFFFFFFFu);        // main.c:206
}

// Row-major over source cells; the k-th cell that needs a marker uses draws 2k and 2k+1 of
// the stream (y jitter first: gcc evaluates v2f's second argument first, see oracle).
__global__ void __launch_bounds__(1024) k_sources(
    Grid g, float h, const unsigned int* __restrict__ cells, size_t ncells,
    uint8_t* __restrict__ count, float2* __restrict__ markers, size_t max_markers,
    const unsigned long long* __restrict__ jump, DevScalars* sc) {
  const unsigned long long n0 = sc->n_markers;
  const unsigned long long cap = max_markers - 1;            // main.c:281
  const bool exhausted0 = sc->source_exhausted || n0 == cap;
  const unsigned long long allow = exhausted0 ? 0 : cap - n0;
  const unsigned long long state0 = sc->rng_state;
  __shared__ unsigned int warp_tot[32];
  __shared__ unsigned long long carry_sh;
  if (threadIdx.x == 0) carry_sh = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (size_t base = 0; base < ncells; base += 1024) {
    const size_t i = base + threadIdx.x;
    bool need = false;
    unsigned int cell = 0;
    if (i < ncells) {
      cell = cells[i];
      need = count[cell] < 4;                                // main.c:287
    }
    const unsigned int bal = __ballot_sync(EULER_FULL_MASK, need);
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    unsigned int woff = 0, total = 0;
    for (int wv = 0; wv < 32; ++wv) {
      const unsigned int t = warp_tot[wv];
      if (wv < wid) woff += t;
      total += t;
    }
    const unsigned long long rank = carry_sh + woff + __popc(bal & ((1u << lane) - 1u));
    if (need && rank < allow) {
      unsigned long long s = rng_jump(jump, state0, 2 * rank);
      s = rng_step(s);
      const float jy = rng_float(s);
      s = rng_step(s);
      const float jx = rng_float(s);
      const int y = (int)(cell / (unsigned int)g.pitch), x = (int)(cell % (unsigned int)g.pitch);
      markers[n0 + rank] = make_float2(h * (x + jx), h * (y + jy));   // main.c:288
      count[cell] += 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_sh += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const unsigned long long added = carry_sh < allow ? carry_sh : allow;
    sc->n_markers = n0 + added;
    sc->rng_state = rng_jump(jump, state0, 2 * added);
    sc->source_exhausted = (exhausted0 || (n0 + added == cap)) ? 1 : 0;
  }
}

}  // namespace

void init_rng_jump_table(unsigned long long* t) {
  // column b of T: image of the basis vector e_b under one xorshift64 step
  for (int b = 0; b < 64; ++b) {
    unsigned long long s = 1ull << b;
    s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
    t[b] = s;
  }
  for (int j = 1; j < 64; ++j) {
    const unsigned long long* prev = t + (size_t)(j - 1) * 64;
    unsigned long long* cur = t + (size_t)j * 64;
    for (int b = 0; b < 64; ++b) {
      unsigned long long s = prev[b], out = 0;
      for (int k = 0; k < 64; ++k)
        if ((s >> k) & 1ull) out ^= prev[k];
      cur[b] = out;
    }
  }
}

void launch_advect_markers(Ctx& c, float dt, int /*mode*/) {
  ProfScope ps(c, KC_ADVECT_MARKERS);
  const int blocks = c.sm_count * 8;
  k_advect_markers<<<blocks, MTHREADS, 0, c.stream>>>(c.g, c.lim, c.u, c.v, c.count, c.solid,
                                                      c.h, c.markers, c.markers_alt, c.sc, dt);
  c.launches += 1;
  float2* t = c.markers; c.markers = c.markers_alt; c.markers_alt = t;
}

void launch_refresh_counts(Ctx& c) {
  ProfScope ps(c, KC_REFRESH_COUNTS);
  // prev <- cur (main.c:103) by swapping planes; cur is rebuilt from scratch (main.c:104)
  uint8_t* t = c.prev_count; c.prev_count = c.count; c.count = t;
  const int blocks = c.sm_count * 8;
  k_bin_markers<<<blocks, MTHREADS, 0, c.stream>>>(c.g, c.h, c.markers, c.sink, c.solid,
                                                   c.count32, c.seg_count, c.sc);
  k_seg_scan<<<1, 1024, 0, c.stream>>>(c.seg_count, c.seg_offset, c.sc);
  unsigned int* del_list = reinterpret_cast<unsigned int*>(c.markers_alt);
  k_list_deleted<<<blocks, MTHREADS, 0, c.stream>>>(c.g, c.h, c.markers, c.sink, c.solid,
                                                    c.seg_count, c.seg_offset, del_list, c.sc);
  k_fill_holes<<<1, 1024, 0, c.stream>>>(c.g, c.h, c.markers, c.sink, c.solid, del_list, c.sc);
  k_fold_counts<<<blocks, 256, 0, c.stream>>>(c.g, c.count32, c.count);
  c.launches += 5;
}

void launch_sources(Ctx& c) {
  if (c.n_source_cells == 0) return;
  ProfScope ps(c, KC_SOURCES);
  k_sources<<<1, 1024, 0, c.stream>>>(c.g, c.h, c.source_cells, c.n_source_cells, c.count,
                                      c.markers, c.max_markers, c.rng_jump, c.sc);
  c.launches += 1;
}

}  // namespace euler
