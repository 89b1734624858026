// tests/csrc/pcg_ops_host.cpp — TEST INFRASTRUCTURE: the row operators of the pressure-solve
// kernels (euler_b200/csrc/pcg_ops.cuh), compiled for the HOST with g++ -ffp-contract=off and
// driven row by row over padded host planes, so that their arithmetic can be compared bit for
// bit with the oracle on a machine without a GPU (tests/test_kernel_arith_host.py).
//
// What runs here is the SAME source the GPU kernels instantiate; what is emulated is only the
// feeding: on the GPU the three input rows of a call are stages of the TMA ring in shared
// memory (pcg_pipe.cuh run()), here the row views point straight into the planes.  Planes have
// the device layout: row-major, `pitch` elements per row (multiple of 32), zero guard rows
// above and below, pointers given for row 0.  "Threads" are visited in row-major order with ONE
// operator instance, so the fused dot products come out as the sequential row-major sums of the
// oracle's dot().
#include <stdint.h>

#include "pcg_ops.cuh"

using namespace euler;

namespace {

Grid make_grid(int nx, int ny, int pitch) {
  Grid g;
  g.nx = nx; g.ny = ny; g.pitch = pitch; g.yoff = 0; g.gny = ny; g.th = 32;
  return g;
}

// calls op.row for every thread position of every row, C cells per thread, like pipe::run
template <int ND, int NB, int C, class T, class Op>
void sweep(const Grid& g, const T* const (&d)[ND], const uint8_t* const (&b)[NB], Op& op) {
  for (int y = 0; y < g.ny; ++y)
    for (int x0 = 0; x0 < g.pitch; x0 += pipe::TW) {
      const int w = g.pitch - x0 < pipe::TW ? g.pitch - x0 : pipe::TW;
      pipe::RowView<ND, NB, T> v[3];
      for (int k = 0; k < 3; ++k) {
        const long row = (long)(y - 1 + k) * g.pitch + x0;
        for (int i = 0; i < ND; ++i) v[k].d[i] = d[i] + row;
        for (int i = 0; i < NB; ++i) v[k].b[i] = b[i] + row;
      }
      for (int t4 = 0; t4 < pipe::TW; t4 += C) op.row(v[0], v[1], v[2], t4, x0 + t4, y, t4 < w);
    }
}

template <int C, class T>
void rb_forward(int nx, int ny, int pitch, const T* r, const T* pc, const uint8_t* fluid, T* q) {
  const Grid g = make_grid(nx, ny, pitch);
  RbForwardPipe<C, T> op{g, q};
  const T* const d[2] = {r, pc};
  const uint8_t* const b[1] = {fluid};
  sweep<2, 1, C, T>(g, d, b, op);
}

template <int C, class T>
double rb_backward(int nx, int ny, int pitch, const T* q, const T* pc, const T* r, const uint8_t* fluid, T* z) {
  const Grid g = make_grid(nx, ny, pitch);
  RbBackwardPipe<C, T> op{g, z, 0.0, 0, ny, nullptr, nullptr, 0, false};
  const T* const d[3] = {q, pc, r};
  const uint8_t* const b[1] = {fluid};
  sweep<3, 1, C, T>(g, d, b, op);
  return op.acc;
}

template <int C, class T>
double fused_search_apply(int nx, int ny, int pitch, const T* z, const T* s, const uint8_t* fluid,
                          const int8_t* adiag, double beta, int init, T* s_new, T* as) {
  const Grid g = make_grid(nx, ny, pitch);
  FusedSearchApply<C, T> op{g, s_new, as, (T)beta, init != 0, 0.0, 0, ny};
  const T* const d[2] = {z, s};
  const uint8_t* const b[2] = {fluid, reinterpret_cast<const uint8_t*>(adiag)};
  sweep<2, 2, C, T>(g, d, b, op);
  return op.acc;
}

}  // namespace

#define DISPATCH_C(call2, call4) (cpt == 2 ? (call2) : (call4))

extern "C" {

void ops_rb_forward_f64(int nx, int ny, int pitch, int cpt, const double* r, const double* pc, const uint8_t* fluid, double* q) {
  if (cpt == 2) rb_forward<2, double>(nx, ny, pitch, r, pc, fluid, q); else rb_forward<4, double>(nx, ny, pitch, r, pc, fluid, q);
}
void ops_rb_forward_f32(int nx, int ny, int pitch, int cpt, const float* r, const float* pc, const uint8_t* fluid, float* q) {
  if (cpt == 2) rb_forward<2, float>(nx, ny, pitch, r, pc, fluid, q); else rb_forward<4, float>(nx, ny, pitch, r, pc, fluid, q);
}
double ops_rb_backward_f64(int nx, int ny, int pitch, int cpt, const double* q, const double* pc, const double* r,
                           const uint8_t* fluid, double* z) {
  return DISPATCH_C((rb_backward<2, double>(nx, ny, pitch, q, pc, r, fluid, z)), (rb_backward<4, double>(nx, ny, pitch, q, pc, r, fluid, z)));
}
double ops_rb_backward_f32(int nx, int ny, int pitch, int cpt, const float* q, const float* pc, const float* r,
                           const uint8_t* fluid, float* z) {
  return DISPATCH_C((rb_backward<2, float>(nx, ny, pitch, q, pc, r, fluid, z)), (rb_backward<4, float>(nx, ny, pitch, q, pc, r, fluid, z)));
}
double ops_fused_search_apply_f64(int nx, int ny, int pitch, int cpt, const double* z, const double* s, const uint8_t* fluid,
                                  const int8_t* adiag, double beta, int init, double* s_new, double* as) {
  return DISPATCH_C((fused_search_apply<2, double>(nx, ny, pitch, z, s, fluid, adiag, beta, init, s_new, as)),
                    (fused_search_apply<4, double>(nx, ny, pitch, z, s, fluid, adiag, beta, init, s_new, as)));
}
double ops_fused_search_apply_f32(int nx, int ny, int pitch, int cpt, const float* z, const float* s, const uint8_t* fluid,
                                  const int8_t* adiag, double beta, int init, float* s_new, float* as) {
  return DISPATCH_C((fused_search_apply<2, float>(nx, ny, pitch, z, s, fluid, adiag, beta, init, s_new, as)),
                    (fused_search_apply<4, float>(nx, ny, pitch, z, s, fluid, adiag, beta, init, s_new, as)));
}
// z = A s (+ z.s), the unfused fp64 operator (k_apply_a_pipe)
double ops_apply_a_f64(int nx, int ny, int pitch, const double* s, const uint8_t* fluid, const int8_t* adiag, double* z) {
  const Grid g = make_grid(nx, ny, pitch);
  ApplyAPipe op{g, z, 0.0, 0, ny};
  const double* const d[1] = {s};
  const uint8_t* const b[2] = {fluid, reinterpret_cast<const uint8_t*>(adiag)};
  sweep<1, 2, 4, double>(g, d, b, op);
  return op.acc;
}
// r32 = fp32(b - A p), the residual replacement of the mixed-precision mode (k_true_residual)
void ops_true_residual(int nx, int ny, int pitch, const double* p, const double* bvec, const uint8_t* fluid,
                       const int8_t* adiag, float* r) {
  const Grid g = make_grid(nx, ny, pitch);
  TrueResidual op{g, bvec, r};
  const double* const d[1] = {p};
  const uint8_t* const b[2] = {fluid, reinterpret_cast<const uint8_t*>(adiag)};
  sweep<1, 2, 4, double>(g, d, b, op);
}

// stencil_variant 2 (kept for A/B): p += alpha s, r' = r - alpha A s, ||r'||inf and q = L^-1 r'
// in one pass (k_fused_axpy_forward); returns ||r'||inf
double ops_fused_axpy_forward_f64(int nx, int ny, int pitch, const double* r, const double* as, const double* pc,
                                  const uint8_t* fluid, const double* s, double alpha, double* p, double* r_new,
                                  double* q) {
  const Grid g = make_grid(nx, ny, pitch);
  FusedAxpyForward<2> op{g, s, p, r_new, q, alpha, 0.0, 0, ny, {}, {}, ~(size_t)0};
  const double* const d[3] = {r, as, pc};
  const uint8_t* const b[1] = {fluid};
  sweep<3, 1, 2, double>(g, d, b, op);
  return op.mx;
}

}  // extern "C"
