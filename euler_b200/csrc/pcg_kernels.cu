// euler_b200/csrc/pcg_kernels.cu — Bridson's PCG for the pressure Poisson system, the part of
// project() at reference main.c:735-767, with the helpers it calls:
//
//   k_apply_a        z = A s, fused with the z.s reduction and alpha      main.c:679-691, 752
//   k_axpy           p += alpha s ; r -= alpha z ; ||r||inf ; stop test   main.c:694-702, 753-758
//   k_update_search  s = z + beta s                                       main.c:669-677, 764
//   k_rb_*           red-black ordered IC(0) preconditioner (GPU-parallel mode; not in the
//                    reference — CPU mirror in oracle/euler_oracle.c precon_redblack)
//
// The reference-faithful natural-order IC(0) lives in wavefront.cu.
//
// Scalars never visit the host: each reducing kernel leaves per-block partials, the block
// that finishes last folds them in a fixed order and writes alpha / beta / sigma / the stop
// flag into DevScalars (common.cuh grid_reduce_last_block).  Once `done` is set every later
// kernel of the solve returns immediately, so the host may enqueue iterations in batches and
// poll rarely; the iterate sequence is the same as with a per-iteration test.
//
// All vectors are fp64 like the reference's; only fluid cells are read or written
// (is_fluid(y,x) guards every loop body of the reference).
#include "kernels.h"

namespace euler {

namespace {

constexpr int BX = 32, BY = 8;
inline dim3 grid2d(const Grid& g) { return dim3((g.nx + BX - 1) / BX, (g.ny + BY - 1) / BY); }

enum { CTR_ZS = 0, CTR_NORM = 1, CTR_ZR = 2 };

// z = A s on fluid cells; off-diagonals are -1 towards fluid neighbours, a_diag = number of
// non-solid neighbours.  Subtraction order as in main.c:683-687: right, up, left, down.
__global__ void __launch_bounds__(BX* BY) k_apply_a(
    Grid g, const double* __restrict__ s, const uint8_t* __restrict__ fluid,
    const int8_t* __restrict__ adiag, double* __restrict__ z, double* partials, DevScalars* sc) {
  if (sc->done) return;
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  double prod = 0.0;
  if (x < g.nx && y < g.ny) {
    const size_t c = gidx(g, x, y);
    if (fluid[c]) {
      const double sc_ = s[c];
      double out = (double)adiag[c] * sc_;
      out -= fluid[c + 1] ? s[c + 1] : 0.0;
      out -= fluid[c + g.pitch] ? s[c + g.pitch] : 0.0;
      out -= fluid[c - 1] ? s[c - 1] : 0.0;
      out -= fluid[c - g.pitch] ? s[c - g.pitch] : 0.0;
      z[c] = out;
      prod = out * sc_;
    }
  }
  const double bsum = block_reduce<false>(prod);
  grid_reduce_last_block<false>(bsum, partials, &sc->ctr[CTR_ZS], [&](double total) {
    sc->zs = total;
    sc->alpha = sc->sigma / total;                           // main.c:752
  });
}

__global__ void __launch_bounds__(BX* BY) k_axpy(
    Grid g, const double* __restrict__ s, const double* __restrict__ z,
    const uint8_t* __restrict__ fluid, double* __restrict__ p, double* __restrict__ r,
    double* partials, DevScalars* sc, double tol) {
  if (sc->done) return;
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  const double alpha = sc->alpha;
  double m = 0.0;
  if (x < g.nx && y < g.ny) {
    const size_t c = gidx(g, x, y);
    if (fluid[c]) {
      p[c] = p[c] + s[c] * alpha;                            // fmadd(s, alpha, p)  main.c:753
      const double rn = r[c] + z[c] * -alpha;                // fmadd(z, -alpha, r) main.c:754
      r[c] = rn;
      m = fabs(rn);
    }
  }
  // NaN-dropping max like `a > maximum` (main.c:659-662)
  m = (m > 0.0) ? m : 0.0;
  const double bmax = block_reduce<true>(m);
  grid_reduce_last_block<true>(bmax, partials, &sc->ctr[CTR_NORM], [&](double total) {
    sc->resid = total;
    sc->iters += 1;
    if (total <= tol) sc->done = 1;                          // main.c:756-758
  });
}

__global__ void __launch_bounds__(BX* BY) k_update_search(
    Grid g, const double* __restrict__ z, const uint8_t* __restrict__ fluid,
    double* __restrict__ s, const DevScalars* sc) {
  if (sc->done) return;
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  if (x >= g.nx || y >= g.ny) return;
  const size_t c = gidx(g, x, y);
  if (fluid[c]) s[c] = z[c] + sc->beta * s[c];               // main.c:673
}

__global__ void __launch_bounds__(BX* BY) k_copy_search(Grid g, const double* __restrict__ z,
                                                        double* __restrict__ s) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  if (x >= g.nx || y >= g.ny) return;
  const size_t c = gidx(g, x, y);
  s[c] = z[c];                                               // memcpy(s, z) main.c:746
}

__global__ void k_pcg_reset(DevScalars* sc) {
  sc->iters = 0; sc->done = 0; sc->resid = 0.0; sc->sigma = 0.0; sc->alpha = 0.0; sc->beta = 0.0;
}

// ----------------------------------------------- red-black ordered IC(0) ----
// Red = (x+y) even, eliminated first.  E_red = a (1 if a==0); E_black = a - sum over fluid
// red neighbours of 1/E_red, with the sigma=0.25 safety rule of main.c:594-596.  Neighbour
// sums always run left, right, down, up.  precon = 1/sqrt(E).

__device__ __forceinline__ double rb_e_red(const int8_t* __restrict__ adiag, size_t c) {
  const double a = (double)adiag[c];
  return a != 0.0 ? a : 1.0;
}

__global__ void __launch_bounds__(BX* BY) k_rb_build(
    Grid g, const uint8_t* __restrict__ fluid, const int8_t* __restrict__ adiag,
    double* __restrict__ precon) {
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  if (x < 1 || y < 1 || x >= g.nx - 1 || y >= g.ny - 1) return;
  const size_t c = gidx(g, x, y);
  if (!fluid[c]) return;
  if (((x + y) & 1) == 0) { precon[c] = 1.0 / sqrt(rb_e_red(adiag, c)); return; }
  const double a = (double)adiag[c];
  double e = a;
  const long off[4] = {-1, 1, -(long)g.pitch, (long)g.pitch};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const size_t nb = c + off[k];
    if (fluid[nb]) e = e - 1.0 / rb_e_red(adiag, nb);
  }
  if (e < 0.25 * a) e = a != 0.0 ? a : 1.0;
  precon[c] = 1.0 / sqrt(e);
}

// q = L^-1 r.  Red: q = r*pc.  Black: q = (r + sum pc_nb * q_nb) * pc with q_nb = r_nb*pc_nb
// recomputed on the fly (identical bits to the stored value).
__global__ void __launch_bounds__(BX* BY) k_rb_forward(
    Grid g, const double* __restrict__ r, const uint8_t* __restrict__ fluid,
    const double* __restrict__ precon, double* __restrict__ q, const DevScalars* sc) {
  if (sc->done) return;
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  if (x < 1 || y < 1 || x >= g.nx - 1 || y >= g.ny - 1) return;
  const size_t c = gidx(g, x, y);
  if (!fluid[c]) return;
  double t = r[c];
  if ((x + y) & 1) {
    const long off[4] = {-1, 1, -(long)g.pitch, (long)g.pitch};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t nb = c + off[k];
      if (fluid[nb]) { const double pn = precon[nb]; t = t + pn * (r[nb] * pn); }
    }
  }
  q[c] = t * precon[c];
}

// z = L^-T q.  Black: z = q*pc.  Red: z = (q + sum pc * z_nb) * pc with z_nb = q_nb*pc_nb.
// Fused with the z.r reduction that follows every preconditioner application
// (main.c:748, 762) and the beta / sigma update (main.c:763-765).
__global__ void __launch_bounds__(BX* BY) k_rb_backward(
    Grid g, const double* __restrict__ q, const double* __restrict__ r,
    const uint8_t* __restrict__ fluid, const double* __restrict__ precon,
    double* __restrict__ z, double* partials, DevScalars* sc, int init) {
  if (sc->done) return;
  const int x = blockIdx.x * BX + threadIdx.x, y = blockIdx.y * BY + threadIdx.y;
  double prod = 0.0;
  if (x >= 1 && y >= 1 && x < g.nx - 1 && y < g.ny - 1) {
    const size_t c = gidx(g, x, y);
    if (fluid[c]) {
      const double pc = precon[c];
      double t = q[c];
      if (((x + y) & 1) == 0) {
        const long off[4] = {-1, 1, -(long)g.pitch, (long)g.pitch};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const size_t nb = c + off[k];
          if (fluid[nb]) t = t + pc * (q[nb] * precon[nb]);
        }
      }
      const double zc = t * pc;
      z[c] = zc;
      prod = zc * r[c];
    }
  }
  const double bsum = block_reduce<false>(prod);
  grid_reduce_last_block<false>(bsum, partials, &sc->ctr[CTR_ZR], [&](double total) {
    if (init) { sc->sigma = total; }                         // main.c:748
    else { sc->beta = total / sc->sigma; sc->sigma = total; }  // main.c:762-765
  });
}

}  // namespace

void launch_pcg_reset(Ctx& c) {
  k_pcg_reset<<<1, 1, 0, c.stream>>>(c.sc);
  c.launches += 1;
}

void launch_apply_a(Ctx& c, bool) {
  k_apply_a<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(c.g, c.s, c.count, c.adiag, c.z,
                                                        c.partials, c.sc);
  c.launches += 1;
}

void launch_axpy(Ctx& c, double tol) {
  k_axpy<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(c.g, c.s, c.z, c.count, c.p, c.r,
                                                     c.partials, c.sc, tol);
  c.launches += 1;
}

void launch_update_search(Ctx& c) {
  k_update_search<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(c.g, c.z, c.count, c.s, c.sc);
  c.launches += 1;
}

void launch_copy_search(Ctx& c) {
  k_copy_search<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(c.g, c.z, c.s);
  c.launches += 1;
}

void launch_rb_build(Ctx& c) {
  k_rb_build<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(c.g, c.count, c.adiag, c.precon);
  c.launches += 1;
}

void launch_rb_apply(Ctx& c, bool init) {
  k_rb_forward<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(c.g, c.r, c.count, c.precon, c.q, c.sc);
  k_rb_backward<<<grid2d(c.g), dim3(BX, BY), 0, c.stream>>>(c.g, c.q, c.r, c.count, c.precon,
                                                            c.z, c.partials, c.sc, init ? 1 : 0);
  c.launches += 2;
}

}  // namespace euler
