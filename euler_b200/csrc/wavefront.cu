// euler_b200/csrc/wavefront.cu — reference-faithful IC(0) ("MIC(0)-style", sigma = 0.25,
// tau = 0) preconditioner of reference main.c:580-627, natural (row-major) ordering.
//
// The three loops of apply_preconditioner — E^-1 build (:586-600), L q = r (:603-613) and
// L^T z = q (:616-626) — carry a sequential dependency on the left/lower (resp. right/upper)
// neighbour, so cell (x,y) can only be done after (x-1,y) and (x,y-1): an anti-diagonal
// wavefront.  Layout of the parallel sweep:
//
//   * the interior rows are cut into STRIPS of 32 rows; one warp owns a strip, lane l owns
//     row y0+l and at step t works on column x = 1 + t - l (lanes are skewed by one column);
//   * the value of the left neighbour never leaves the lane's registers; the value of the
//     lower neighbour is what lane l-1 produced one step earlier -> one __shfl_up per step;
//   * lane 0 needs the top row of the strip below.  Strips are chained through a progress
//     counter per strip (st.release / ld.acquire, published every 8 columns), so
//     strip k runs ~32+8 columns behind strip k-1: a software pipeline of warps
//     across the SMs.  Strip ids are handed out by an atomic ticket so a strip can only wait
//     on a strip that has already started (no reliance on block scheduling order).
//   * the backward solve is the mirror image (rows and columns descending).
//
// Arithmetic is the reference's, operation for operation (fp64, no FMA), so q, z and
// g_precon are bit-identical to the CPU's given the same r — including quirk SURVEY §9.1:
// the diagonal recurrence reads precon of the left/lower neighbour even when that cell is
// not fluid (a value left over from an earlier time step; precon is only written at fluid
// cells and never cleared).
//
// Byte traffic is the same 56 B/cell as a streaming kernel would need; the sweep is bound
// by the dependent fp64 chain (mul, add, add, mul per column), not by HBM — see DESIGN.md.
#include "kernels.h"

#include <stdlib.h>

namespace euler {

namespace {

constexpr int WF_PREFETCH = 32;   // columns ahead for prefetch.global.L1

enum { WF_BUILD = 0, WF_FORWARD = 1, WF_BACKWARD = 2 };

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// MODE WF_BUILD   : precon(x,y) from adiag and precon(x-1,y), precon(x,y-1)   main.c:586-600
// MODE WF_FORWARD : q = L^-1 r                                                main.c:603-613
// MODE WF_BACKWARD: z = L^-T q, fused with the z.r reduction                  main.c:616-626
//
// `in`  = r (forward) / q (backward); `out` = q (forward) / z (backward).
template <int MODE>
__global__ void __launch_bounds__(32) k_ic0_sweep(
    Grid g, const uint8_t* __restrict__ fluid, const int8_t* __restrict__ adiag,
    double* precon, const double* __restrict__ in, double* out, const double* __restrict__ rvec,
    unsigned int* progress, unsigned int* ticket, double* partials, DevScalars* sc, int init,
    int exact, int dbg, int start_lag) {
  if (MODE != WF_BUILD && sc->done) return;
  const int lane = threadIdx.x;
  const int n_strips = gridDim.x;
  unsigned int strip = 0;
  if (lane == 0) strip = atomicAdd(ticket, 1u) % (unsigned int)n_strips;
  strip = __shfl_sync(EULER_FULL_MASK, strip, 0);

  const int ncols = g.nx - 2;                       // interior columns 1 .. nx-2
  const bool fwd = MODE != WF_BACKWARD;
  // forward: rows ascend from 1; backward: rows descend from ny-2
  const int y = fwd ? 1 + 32 * (int)strip + lane : (g.ny - 2) - 32 * (int)strip - lane;
  const bool row_ok = y >= 1 && y <= g.ny - 2;
  const int ystep = fwd ? 1 : -1;                   // direction towards "later" rows
  const int y_dep = y - ystep;                      // row this row depends on
  // the lane that owns the strip's last row publishes progress for the next strip
  const bool publisher = lane == 31;

  const size_t row = (size_t)(row_ok ? y : 1) * g.pitch;
  const size_t row_dep = (size_t)(row_ok ? y_dep : 1) * g.pitch;

  // carried along the row (value at the previous column of this lane's row)
  double prev_val = 0.0;                            // q(x-1) / z(x+1); 0 outside (memset)
  double prev_pc = 0.0;                             // precon(x-1) for build/forward
  bool prev_fl = false;                             // fluid(x+1) for backward
  if (row_ok && fwd) prev_pc = precon[row + 0];     // column 0: never fluid, plane value
  // what this lane produced in the previous step (consumed by lane+1 now)
  double last_val = 0.0, last_pc = 0.0;
  bool last_fl = false;

  unsigned int avail = 0;                           // columns of the dependency row known done
  double acc = 0.0;

  // Start `start_lag` columns behind the strip below.  All strips advance at the same pace (the
  // dependent fp64 chain), so the distance persists, and with it the one-block-ahead loads of
  // the dependency row (load_dep, non-blocking) always find their columns published: the
  // release -> acquire -> ld.cg hand-off latency (~1-2 us through L2) is paid once per strip
  // instead of once per 8 columns.  Without the head start every strip ran at the minimum
  // distance and blocked on the hand-off in every block — that, not the fp64 chain, bounded the
  // sweep.  Costs n_strips x start_lag extra column steps of pipeline fill.
  if (lane == 0 && row_ok && strip > 0 && start_lag > 0 && !(dbg & 1)) {
    const unsigned int want = (unsigned int)min(ncols, start_lag);
    do { avail = ld_acquire(progress + (strip - 1)); } while (avail < want);
  }
  __syncwarp();

  // Steps are processed in blocks of WB columns, software-pipelined one block deep: while block
  // b is being computed out of registers, the loads of block b+1 are already in flight (own
  // row: operand, precon, fluid; lane 0: the row of the strip below, if that strip has
  // published it — otherwise lane 0 falls back to a blocking wait at the top of b+1).  One
  // memory latency is hidden behind WB dependent steps; progress is published once per block.
  constexpr int WB = 8;
  const int nsteps = ncols + 31;
  struct Buf {
    double in_v[WB], pc_v[WB], r_v[WB], dval[WB], dpc[WB];
    unsigned char fl_v[WB], dfl[WB];
    signed char ad_v[WB];
    bool dep_loaded;
  };
  auto load_own = [&](Buf& b, int t0) {
#pragma unroll
    for (int j = 0; j < WB; ++j) {
      const int k = t0 + j - lane;
      const int x = fwd ? 1 + k : (g.nx - 2) - k;
      const bool ok = row_ok && k >= 0 && k < ncols;
      b.in_v[j] = 0.0; b.pc_v[j] = 0.0; b.r_v[j] = 0.0; b.fl_v[j] = 0; b.ad_v[j] = 0;
      if (ok) {
        const size_t c = row + x;
        b.fl_v[j] = fluid[c];
        if (MODE == WF_BUILD) { b.ad_v[j] = adiag[c]; if (!b.fl_v[j]) b.pc_v[j] = precon[c]; }
        else { b.pc_v[j] = precon[c]; b.in_v[j] = in[c]; }
        if (MODE == WF_BACKWARD) b.r_v[j] = rvec[c];
      }
    }
  };
  // lane 0: the dependency row for the block starting at t0; `blocking` = must succeed now
  auto load_dep = [&](Buf& b, int t0, bool blocking) {
    b.dep_loaded = false;
    if (lane != 0 || !row_ok || t0 >= nsteps) return;
    const int klast = min(t0 + WB - 1, ncols - 1);
    if (!(dbg & 1) && strip > 0 && klast >= 0 && avail < (unsigned int)(klast + 1)) {
      avail = ld_acquire(progress + (strip - 1));
      if (avail < (unsigned int)(klast + 1)) {
        if (!blocking) return;
        do { avail = ld_acquire(progress + (strip - 1)); } while (avail < (unsigned int)(klast + 1));
      }
    }
#pragma unroll
    for (int j = 0; j < WB; ++j) {
      const int k = t0 + j;
      const int x = fwd ? 1 + k : (g.nx - 2) - k;
      b.dval[j] = 0.0; b.dpc[j] = 0.0; b.dfl[j] = 0;
      if (k < ncols && !(dbg & 4)) {
        if (strip > 0) {
          if (MODE != WF_BUILD) b.dval[j] = __ldcg(out + row_dep + x);
          b.dpc[j] = __ldcg(precon + row_dep + x);
        } else {
          b.dpc[j] = precon[row_dep + x];                    // border row: never fluid
        }
        b.dfl[j] = fluid[row_dep + x];
      }
    }
    b.dep_loaded = true;
  };

  Buf nxt;
  load_own(nxt, 0);
  load_dep(nxt, 0, true);
  for (int t0 = 0; t0 < nsteps; t0 += WB) {
    Buf cur = nxt;
    if (lane == 0 && row_ok && !cur.dep_loaded) load_dep(cur, t0, true);
    if (t0 + WB < nsteps) {
      load_own(nxt, t0 + WB);
      load_dep(nxt, t0 + WB, false);
    }
    {   // pull the lines of the blocks after next into L1
      const int k = t0 + 2 * WB - lane;
      const int xp = fwd ? 1 + k + WF_PREFETCH : (g.nx - 2) - k - WF_PREFETCH;
      if (row_ok && xp >= 0 && xp < g.nx) {
        if (MODE != WF_BUILD) prefetch_l1(in + row + xp);
        prefetch_l1(precon + row + xp);
        if (MODE == WF_BACKWARD) prefetch_l1(rvec + row + xp);
      }
    }
#pragma unroll
    for (int j = 0; j < WB; ++j) {
      const int k = t0 + j - lane;                  // 0-based position along the sweep
      const int x = fwd ? 1 + k : (g.nx - 2) - k;
      const bool col_ok = k >= 0 && k < ncols;

      double dep_val = __shfl_up_sync(EULER_FULL_MASK, last_val, 1);
      double dep_pc = __shfl_up_sync(EULER_FULL_MASK, last_pc, 1);
      bool dep_fl = __shfl_up_sync(EULER_FULL_MASK, (int)last_fl, 1) != 0;
      if (lane == 0) { dep_val = cur.dval[j]; dep_pc = cur.dpc[j]; dep_fl = cur.dfl[j] != 0; }

      if (col_ok && row_ok) {
        const size_t c = row + x;
        const bool fl = cur.fl_v[j] != 0;
        if (MODE == WF_BUILD) {
          double pc;
          if (fl) {
            const double a = (double)cur.ad_v[j];
            const double wl = -1.0 * prev_pc;                  // get_a_minus_i == -1
            const double wd = -1.0 * dep_pc;
            double e = a - wl * wl - wd * wd;                  // main.c:590-593
            if (e < 0.25 * a) e = a != 0.0 ? a : 1.0;          // main.c:594-596
            pc = 1.0 / sqrt(e);
            precon[c] = pc;
          } else {
            pc = cur.pc_v[j];                                      // stale value stays
          }
          prev_pc = pc;
          last_pc = pc;
        } else if (MODE == WF_FORWARD) {
          const double pc = cur.pc_v[j];
          double qv = 0.0;
          if (fl) {
            const double tt = cur.in_v[j] - (-1.0 * prev_pc) * prev_val - (-1.0 * dep_pc) * dep_val;
            qv = tt * pc;                                      // main.c:607-610
          }
          out[c] = qv;
          prev_val = qv; prev_pc = pc;
          last_val = qv; last_pc = pc;
        } else {
          double zv = 0.0;
          if (fl) {
            const double pc = cur.pc_v[j];
            const double ar = prev_fl ? -1.0 : 0.0;            // get_a_plus_i(y,x)
            const double au = dep_fl ? -1.0 : 0.0;             // get_a_plus_j(y,x)
            const double tt = cur.in_v[j] - ar * pc * prev_val - au * pc * dep_val;
            zv = tt * pc;                                      // main.c:620-623
            acc += zv * cur.r_v[j];
          }
          out[c] = zv;
          prev_val = zv; prev_fl = fl;
          last_val = zv; last_fl = fl;
        }
      }
    }
    if (publisher && row_ok) {
      const int kdone = min(t0 + WB - 1 - lane, ncols - 1);  // last column this lane finished
      if (kdone >= 0 && !(dbg & 2)) st_release(progress + strip, (unsigned int)(kdone + 1));
    }
  }

  if (MODE == WF_BACKWARD) {
    acc = warp_sum(acc);
    grid_reduce_last_block<false>(acc, partials, &sc->ctr[3], [&](double total) {
      if (exact) return;                                     // k_dot_seq supplies z.r instead
      if (init) { sc->sigma = total; }                       // main.c:748
      else { sc->beta = total / sc->sigma; sc->sigma = total; }   // main.c:762-765
    });
  }
}

}  // namespace

static int wf_dbg() { static int v = -1; if (v < 0) { const char* e = getenv("EULER_WF_DEBUG"); v = e ? atoi(e) : 0; } return v; }

// head start of a strip over the next one, in columns (see k_ic0_sweep); EULER_WF_LAG overrides
static int wf_lag() { static int v = -1; if (v < 0) { const char* e = getenv("EULER_WF_LAG"); v = e ? atoi(e) : 64; } return v; }

// progress flags AND the strip tickets restart at 0 for every sweep: a ticket that kept counting
// would lose its alignment with n_strips when it wraps at 2^32 (n_strips is not a power of two),
// and a strip could then be handed out before the one below it
static void reset_wavefront(Ctx& c) {
  cudaMemsetAsync(c.wf_progress, 0, sizeof(unsigned int) * (size_t)c.n_strips, c.stream);
  cudaMemsetAsync(c.sc->ticket, 0, sizeof(c.sc->ticket), c.stream);
}

void launch_ic0_build(Ctx& c) {
  ProfScope ps(c, KC_PRECON_BUILD);
  reset_wavefront(c);
  k_ic0_sweep<WF_BUILD><<<c.n_strips, 32, 0, c.stream>>>(
      c.g, c.count, c.adiag, c.precon, nullptr, nullptr, nullptr, c.wf_progress,
      &c.sc->ticket[0], c.partials, c.sc, 0, 0, wf_dbg(), wf_lag());
  c.launches += 1;
}

void launch_ic0_apply(Ctx& c, bool init) {
  ProfScope ps(c, KC_PRECON_APPLY);
  reset_wavefront(c);
  k_ic0_sweep<WF_FORWARD><<<c.n_strips, 32, 0, c.stream>>>(
      c.g, c.count, c.adiag, c.precon, c.r, c.q, nullptr, c.wf_progress, &c.sc->ticket[1],
      c.partials, c.sc, 0, 0, wf_dbg(), wf_lag());
  reset_wavefront(c);
  k_ic0_sweep<WF_BACKWARD><<<c.n_strips, 32, 0, c.stream>>>(
      c.g, c.count, c.adiag, c.precon, c.q, c.z, c.r, c.wf_progress, &c.sc->ticket[2],
      c.partials, c.sc, init ? 1 : 0, c.dot_mode, wf_dbg(), wf_lag());
  c.launches += 2;
  launch_dot_zr_exact(c, init);
}

}  // namespace euler
