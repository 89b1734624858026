// euler_b200/csrc/pcg_ops.cuh — the per-row arithmetic of the pressure-solve kernels
// (pcg_kernels.cu), separated from the machinery that feeds it.
//
// Every `Op::row(dn, ce, up, t4, x, y, live)` below computes one output row segment of one
// thread from three input rows (pipe::RowView: pointers indexed by tile column, valid for
// [-HX, TW+HX)).  On the GPU the rows are stages of the TMA ring in shared memory
// (pcg_pipe.cuh run()); the same code compiles for the HOST (tests/csrc/pcg_ops_host.cpp,
// g++ -ffp-contract=off), where the row views point straight into padded host planes — that is
// how the arithmetic of these kernels is checked bit for bit against the oracle on a machine
// without a GPU (tests/test_kernel_arith_host.py).  Reference lines: main.c:669-691 (update_search,
// apply_a), :694-702/:753-754 (fmadd), the red-black IC(0) of oracle/euler_oracle.c
// precon_redblack (not in the reference).
#pragma once
#include <math.h>

#include "common.cuh"
#include "pcg_pipe.cuh"

namespace euler {

namespace {

struct D4 { double v[4]; };

__device__ __forceinline__ D4 ld4(const double* __restrict__ p) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  D4 r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
  return r;
}
__device__ __forceinline__ void st4(double* __restrict__ p, const D4& d) {
  *reinterpret_cast<double2*>(p) = make_double2(d.v[0], d.v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(d.v[2], d.v[3]);
}
struct F4 { float v[4]; };
__device__ __forceinline__ F4 ld4(const float* __restrict__ p) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  F4 r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  return r;
}
__device__ __forceinline__ void st4(float* __restrict__ p, const F4& d) {
  *reinterpret_cast<float4*>(p) = make_float4(d.v[0], d.v[1], d.v[2], d.v[3]);
}
template <class T> struct Vec4;
template <> struct Vec4<double> { using type = D4; };
template <> struct Vec4<float> { using type = F4; };
__device__ __forceinline__ double abs_of(double a) { return fabs(a); }
__device__ __forceinline__ float abs_of(float a) { return fabsf(a); }

__device__ __forceinline__ unsigned ldmask(const uint8_t* __restrict__ p) {
  return *reinterpret_cast<const unsigned*>(p);            // 4 cells, one byte each
}
__device__ __forceinline__ bool mbit(unsigned m, int k) { return ((m >> (8 * k)) & 0xffu) != 0; }

__device__ __forceinline__ unsigned lds_mask4(const uint8_t* p) {       // 4-aligned
  return *reinterpret_cast<const unsigned*>(p);
}
__device__ __forceinline__ D4 lds4(const double* p) {                   // 16 B aligned
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  D4 r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
  return r;
}

template <int C, class T = double> struct DV { T v[C]; };
template <int C> __device__ __forceinline__ DV<C> ldsv(const double* p) {      // 16 B aligned
  DV<C> r;
#pragma unroll
  for (int i = 0; i < C; i += 2) {
    const double2 a = *reinterpret_cast<const double2*>(p + i);
    r.v[i] = a.x; r.v[i + 1] = a.y;
  }
  return r;
}
template <int C> __device__ __forceinline__ DV<C, float> ldsv(const float* p) {   // C*4 B aligned
  DV<C, float> r;
  if (C == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    r.v[0] = a.x; r.v[1] = a.y; r.v[C - 2] = a.z; r.v[C - 1] = a.w;
  } else {
    const float2 a = *reinterpret_cast<const float2*>(p);
    r.v[0] = a.x; r.v[1] = a.y;
  }
  return r;
}
template <int C> __device__ __forceinline__ void stv(double* __restrict__ p, const DV<C>& d) {
#pragma unroll
  for (int i = 0; i < C; i += 2) *reinterpret_cast<double2*>(p + i) = make_double2(d.v[i], d.v[i + 1]);
}
template <int C> __device__ __forceinline__ void stv(float* __restrict__ p, const DV<C, float>& d) {
  if (C == 4) *reinterpret_cast<float4*>(p) = make_float4(d.v[0], d.v[1], d.v[C - 2], d.v[C - 1]);
  else *reinterpret_cast<float2*>(p) = make_float2(d.v[0], d.v[1]);
}
template <int C> __device__ __forceinline__ DV<C> ldg_v(const double* __restrict__ p) { return ldsv<C>(p); }
template <int C> __device__ __forceinline__ unsigned ldsm(const uint8_t* p) {   // C-byte aligned
  if (C == 4) return *reinterpret_cast<const unsigned*>(p);
  return (unsigned)*reinterpret_cast<const unsigned short*>(p);
}

struct ApplyAPipe {
  // planes: d0 = s ; b0 = fluid, b1 = adiag
  const Grid g;
  double* __restrict__ z;
  double acc;
  int a0, a1;
  __device__ __forceinline__ void row(const pipe::RowView<1, 2>& dn, const pipe::RowView<1, 2>& ce,
                                      const pipe::RowView<1, 2>& up, int t4, int x, int y, bool live) {
    const unsigned mc = live ? lds_mask4(ce.b[0] + t4) : 0u;
    if (!mc) return;
    const unsigned md = lds_mask4(dn.b[0] + t4), mu = lds_mask4(up.b[0] + t4);
    const bool fl = ce.b[0][t4 - 1] != 0, fr = ce.b[0][t4 + 4] != 0;
    const unsigned am = lds_mask4(ce.b[1] + t4);
    const D4 sc = lds4(ce.d[0] + t4), sd = lds4(dn.d[0] + t4), su = lds4(up.d[0] + t4);
    const double sl = ce.d[0][t4 - 1], sr = ce.d[0][t4 + 4];
    D4 out;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      out.v[k] = 0.0;
      if (!mbit(mc, k)) continue;
      // main.c:683-687: a_diag*s - right - up - left - down, each only towards fluid
      double o = (double)(int)(signed char)((am >> (8 * k)) & 0xffu) * sc.v[k];
      const bool r_ok = k == 3 ? fr : mbit(mc, k + 1);
      const bool l_ok = k == 0 ? fl : mbit(mc, k - 1);
      o -= r_ok ? (k == 3 ? sr : sc.v[k + 1]) : 0.0;
      o -= mbit(mu, k) ? su.v[k] : 0.0;
      o -= l_ok ? (k == 0 ? sl : sc.v[k - 1]) : 0.0;
      o -= mbit(md, k) ? sd.v[k] : 0.0;
      out.v[k] = o;
      if (y >= a0 && y < a1) acc += o * sc.v[k];
    }
    st4(z + gidx(g, x, y), out);
  }
};

template <int C, class T = double>
struct RbForwardPipe {
  // planes: d0 = r, d1 = pc ; b0 = fluid
  using RV = pipe::RowView<2, 1, T>;
  const Grid g;
  T* __restrict__ q;
  __device__ __forceinline__ static T w(T r, T p) { return p * (r * p); }
  __device__ __forceinline__ void row(const RV& dn, const RV& ce, const RV& up, int t4, int x, int y, bool live) {
    const unsigned mc = live ? ldsm<C>(ce.b[0] + t4) : 0u;
    if (!mc) return;
    const unsigned md = ldsm<C>(dn.b[0] + t4), mu = ldsm<C>(up.b[0] + t4);
    const bool fl = ce.b[0][t4 - 1] != 0, fr = ce.b[0][t4 + C] != 0;
    const DV<C, T> rc = ldsv<C>(ce.d[0] + t4), pc = ldsv<C>(ce.d[1] + t4);
    const DV<C, T> rd = ldsv<C>(dn.d[0] + t4), pd = ldsv<C>(dn.d[1] + t4);
    const DV<C, T> ru = ldsv<C>(up.d[0] + t4), pu = ldsv<C>(up.d[1] + t4);
    const T wl = w(ce.d[0][t4 - 1], ce.d[1][t4 - 1]), wr = w(ce.d[0][t4 + C], ce.d[1][t4 + C]);
    DV<C, T> out;
#pragma unroll
    for (int k = 0; k < C; ++k) {
      out.v[k] = (T)0;
      if (!mbit(mc, k)) continue;
      T t = rc.v[k];
      if ((x + k + y + g.yoff) & 1) {                        // black: + sum over red neighbours
        const bool l_ok = k == 0 ? fl : mbit(mc, k - 1);
        const bool r_ok = k == C - 1 ? fr : mbit(mc, k + 1);
        if (l_ok) t = t + (k == 0 ? wl : w(rc.v[k - 1], pc.v[k - 1]));
        if (r_ok) t = t + (k == C - 1 ? wr : w(rc.v[k + 1], pc.v[k + 1]));
        if (mbit(md, k)) t = t + w(rd.v[k], pd.v[k]);
        if (mbit(mu, k)) t = t + w(ru.v[k], pu.v[k]);
      }
      out.v[k] = t * pc.v[k];
    }
    stv<C>(q + gidx(g, x, y), out);
  }
};

template <int C, class T = double>
struct RbBackwardPipe {
  // planes: d0 = q, d1 = pc, d2 = r ; b0 = fluid
  using RV = pipe::RowView<3, 1, T>;
  const Grid g;
  T* __restrict__ z;
  double acc;
  int a0, a1;
  // slab mode, NVLink path: the neighbours' z planes (biased, see DistArgs) and the halo depth
  T* __restrict__ z_dn;
  T* __restrict__ z_up;
  int depth;
  bool peer_stored;
  __device__ __forceinline__ void row(const RV& dn, const RV& ce, const RV& up, int t4, int x, int y, bool live) {
    const unsigned mc = live ? ldsm<C>(ce.b[0] + t4) : 0u;
    if (!mc) return;
    const unsigned md = ldsm<C>(dn.b[0] + t4), mu = ldsm<C>(up.b[0] + t4);
    const bool fl = ce.b[0][t4 - 1] != 0, fr = ce.b[0][t4 + C] != 0;
    const DV<C, T> qc = ldsv<C>(ce.d[0] + t4), pc = ldsv<C>(ce.d[1] + t4), rc = ldsv<C>(ce.d[2] + t4);
    const DV<C, T> qd = ldsv<C>(dn.d[0] + t4), pd = ldsv<C>(dn.d[1] + t4);
    const DV<C, T> qu = ldsv<C>(up.d[0] + t4), pu = ldsv<C>(up.d[1] + t4);
    const T zl = ce.d[0][t4 - 1] * ce.d[1][t4 - 1], zr = ce.d[0][t4 + C] * ce.d[1][t4 + C];
    DV<C, T> out;
#pragma unroll
    for (int k = 0; k < C; ++k) {
      out.v[k] = (T)0;
      if (!mbit(mc, k)) continue;
      T zc;
      const T p = pc.v[k];
      if ((x + k + y + g.yoff) & 1) {
        zc = qc.v[k] * p;                                    // black: q*pc
      } else {
        const bool l_ok = k == 0 ? fl : mbit(mc, k - 1);
        const bool r_ok = k == C - 1 ? fr : mbit(mc, k + 1);
        T t = qc.v[k];
        if (l_ok) t = t + p * (k == 0 ? zl : qc.v[k - 1] * pc.v[k - 1]);
        if (r_ok) t = t + p * (k == C - 1 ? zr : qc.v[k + 1] * pc.v[k + 1]);
        if (mbit(md, k)) t = t + p * (qd.v[k] * pd.v[k]);
        if (mbit(mu, k)) t = t + p * (qu.v[k] * pu.v[k]);
        zc = t * p;
      }
      out.v[k] = zc;
      if (y >= a0 && y < a1) acc += (double)zc * (double)rc.v[k];
    }
    // only the owned rows are stored: the halo rows of z belong to the neighbouring slabs,
    // which write them directly (NVLink peer stores) or through the halo exchange
    if (y >= a0 && y < a1) {
      const size_t c = gidx(g, x, y);
      stv<C>(z + c, out);
      // my edge rows are the neighbours' halo rows: stored there as they are produced
      if (z_dn && y < a0 + depth) { stv<C>(z_dn + c, out); peer_stored = true; }
      if (z_up && y >= a1 - depth) { stv<C>(z_up + c, out); peer_stored = true; }
    }
  }
};

template <int C, class T = double>
struct FusedSearchApply {
  // planes: d0 = z (M^-1 r), d1 = s ; b0 = fluid, b1 = adiag
  using RV = pipe::RowView<2, 2, T>;
  const Grid g;
  T* __restrict__ s_new;
  T* __restrict__ as;
  T beta;
  bool init;                      // first iteration: s' = z (memcpy(s, z), main.c:746)
  double acc;
  int a0, a1;
  __device__ __forceinline__ T sn(T z, T s) const { return init ? z : z + beta * s; }
  __device__ __forceinline__ void row(const RV& dn, const RV& ce, const RV& up, int t4, int x, int y, bool live) {
    const unsigned mc = live ? ldsm<C>(ce.b[0] + t4) : 0u;
    if (!mc) return;
    const unsigned md = ldsm<C>(dn.b[0] + t4), mu = ldsm<C>(up.b[0] + t4);
    const bool fl = ce.b[0][t4 - 1] != 0, fr = ce.b[0][t4 + C] != 0;
    const unsigned am = ldsm<C>(ce.b[1] + t4);
    const DV<C, T> zc = ldsv<C>(ce.d[0] + t4), sc = ldsv<C>(ce.d[1] + t4);
    const DV<C, T> zd = ldsv<C>(dn.d[0] + t4), sd = ldsv<C>(dn.d[1] + t4);
    const DV<C, T> zu = ldsv<C>(up.d[0] + t4), su = ldsv<C>(up.d[1] + t4);
    const T nl = sn(ce.d[0][t4 - 1], ce.d[1][t4 - 1]), nr = sn(ce.d[0][t4 + C], ce.d[1][t4 + C]);
    DV<C, T> nc, out;
#pragma unroll
    for (int k = 0; k < C; ++k) nc.v[k] = sn(zc.v[k], sc.v[k]);
#pragma unroll
    for (int k = 0; k < C; ++k) {
      out.v[k] = (T)0;
      if (!mbit(mc, k)) { nc.v[k] = sc.v[k]; continue; }       // non-fluid: s untouched
      T o = (T)(int)(signed char)((am >> (8 * k)) & 0xffu) * nc.v[k];
      const bool r_ok = k == C - 1 ? fr : mbit(mc, k + 1);
      const bool l_ok = k == 0 ? fl : mbit(mc, k - 1);
      o -= r_ok ? (k == C - 1 ? nr : sn(zc.v[(k + 1) % C], sc.v[(k + 1) % C])) : (T)0;
      o -= mbit(mu, k) ? sn(zu.v[k], su.v[k]) : (T)0;
      o -= l_ok ? (k == 0 ? nl : sn(zc.v[(k + C - 1) % C], sc.v[(k + C - 1) % C])) : (T)0;
      o -= mbit(md, k) ? sn(zd.v[k], sd.v[k]) : (T)0;
      out.v[k] = o;
      if (y >= a0 && y < a1) acc += (double)o * (double)nc.v[k];
    }
    const size_t c = gidx(g, x, y);
    stv<C>(s_new + c, nc);
    stv<C>(as + c, out);
  }
};

// ---- mixed-precision mode: residual replacement ---------------------------------------------
// r <- b - A p with A p and the subtraction in fp64 (main.c:683-687 applied to p), narrowed to
// fp32 on store.  The fp32 recurrence r -= alpha A s drifts away from the true residual by
// ~2^-24 |A||p| per update; replacing it every `pcg_refresh_every` iterations keeps the
// converged pressure within fp32 rounding of the fp64 solve's (oracle: true_residual32).
// p must be complete (even iteration, see k_axpy).  22 B/cell, once every R iterations.
struct TrueResidual {
  // planes: d0 = p ; b0 = fluid, b1 = adiag
  using RV = pipe::RowView<1, 2>;
  const Grid g;
  const double* __restrict__ b;
  float* __restrict__ r;
  __device__ __forceinline__ void row(const RV& dn, const RV& ce, const RV& up, int t4, int x, int y, bool live) {
    const unsigned mc = live ? lds_mask4(ce.b[0] + t4) : 0u;
    if (!mc) return;
    const unsigned md = lds_mask4(dn.b[0] + t4), mu = lds_mask4(up.b[0] + t4);
    const bool fl = ce.b[0][t4 - 1] != 0, fr = ce.b[0][t4 + 4] != 0;
    const unsigned am = lds_mask4(ce.b[1] + t4);
    const D4 pc = lds4(ce.d[0] + t4), pd = lds4(dn.d[0] + t4), pu = lds4(up.d[0] + t4);
    const double pl = ce.d[0][t4 - 1], pr = ce.d[0][t4 + 4];
    const size_t c = gidx(g, x, y);
    const D4 bv = ld4(b + c);
    F4 out;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      out.v[k] = 0.f;
      if (!mbit(mc, k)) continue;
      double o = (double)(int)(signed char)((am >> (8 * k)) & 0xffu) * pc.v[k];
      const bool r_ok = k == 3 ? fr : mbit(mc, k + 1);
      const bool l_ok = k == 0 ? fl : mbit(mc, k - 1);
      o -= r_ok ? (k == 3 ? pr : pc.v[k + 1]) : 0.0;
      o -= mbit(mu, k) ? pu.v[k] : 0.0;
      o -= l_ok ? (k == 0 ? pl : pc.v[k - 1]) : 0.0;
      o -= mbit(md, k) ? pd.v[k] : 0.0;
      out.v[k] = (float)(bv.v[k] - o);
    }
    st4(r + c, out);
  }
};

template <int C>
struct FusedAxpyForward {
  // planes: d0 = r, d1 = A s, d2 = pc ; b0 = fluid
  const Grid g;
  const double* __restrict__ s;
  double* __restrict__ p;
  double* __restrict__ r_new;
  double* __restrict__ q;
  double alpha;
  double mx;
  int a0, a1;
  // s and p are plain element-wise operands (no halo): they bypass the TMA ring and are
  // fetched one row ahead into registers instead
  DV<C> s_next, p_next;
  size_t c_next;
  using RV = pipe::RowView<3, 1>;
  __device__ __forceinline__ double rn(double r, double as) const {
    return r + as * -alpha;                                  // fmadd(z, -alpha, r), main.c:754
  }
  // pc * (r' * pc): what a RED neighbour contributes to a black cell's forward solve
  __device__ __forceinline__ double wred(double r, double as, double pc) const { return pc * (rn(r, as) * pc); }
  __device__ __forceinline__ void row(const RV& dn, const RV& ce, const RV& up, int t4, int x, int y,
                                      bool live) {
    const unsigned mc = live ? ldsm<C>(ce.b[0] + t4) : 0u;
    if (!mc) return;
    const unsigned md = ldsm<C>(dn.b[0] + t4), mu = ldsm<C>(up.b[0] + t4);
    const bool fl = ce.b[0][t4 - 1] != 0, fr = ce.b[0][t4 + C] != 0;
    const DV<C> rc = ldsv<C>(ce.d[0] + t4), ac = ldsv<C>(ce.d[1] + t4), pc = ldsv<C>(ce.d[2] + t4);
    const DV<C> rd = ldsv<C>(dn.d[0] + t4), ad = ldsv<C>(dn.d[1] + t4), pd = ldsv<C>(dn.d[2] + t4);
    const DV<C> ru = ldsv<C>(up.d[0] + t4), au = ldsv<C>(up.d[1] + t4), pu = ldsv<C>(up.d[2] + t4);
    const double wl = wred(ce.d[0][t4 - 1], ce.d[1][t4 - 1], ce.d[2][t4 - 1]);
    const double wr = wred(ce.d[0][t4 + C], ce.d[1][t4 + C], ce.d[2][t4 + C]);
    const int gy = y + g.yoff;
    const size_t c = gidx(g, x, y);
    DV<C> sv, pv;
    if (c_next == c) { sv = s_next; pv = p_next; }
    else { sv = ldg_v<C>(s + c); pv = ldg_v<C>(p + c); }
    c_next = c + g.pitch;                                    // next row of the tile (or a guard
    s_next = ldg_v<C>(s + c_next);                                // row / the next tile's halo: unused)
    p_next = ldg_v<C>(p + c_next);
    DV<C> rout, qout;
#pragma unroll
    for (int k = 0; k < C; ++k) {
      rout.v[k] = rc.v[k];
      qout.v[k] = 0.0;
      if (!mbit(mc, k)) continue;
      pv.v[k] = pv.v[k] + sv.v[k] * alpha;                   // fmadd(s, alpha, p), main.c:753
      const double r1 = rn(rc.v[k], ac.v[k]);
      rout.v[k] = r1;
      double t = r1;
      if ((x + k + gy) & 1) {                                // black: + red neighbours, l r d u
        const bool l_ok = k == 0 ? fl : mbit(mc, k - 1);
        const bool r_ok = k == C - 1 ? fr : mbit(mc, k + 1);
        if (l_ok) t = t + (k == 0 ? wl : wred(rc.v[(k + C - 1) % C], ac.v[(k + C - 1) % C], pc.v[(k + C - 1) % C]));
        if (r_ok) t = t + (k == C - 1 ? wr : wred(rc.v[(k + 1) % C], ac.v[(k + 1) % C], pc.v[(k + 1) % C]));
        if (mbit(md, k)) t = t + wred(rd.v[k], ad.v[k], pd.v[k]);
        if (mbit(mu, k)) t = t + wred(ru.v[k], au.v[k], pu.v[k]);
      }
      qout.v[k] = t * pc.v[k];
      if (y >= a0 && y < a1) {
        const double a = fabs(r1);
        if (a > mx) mx = a;                                  // NaN-dropping max, main.c:659-662
      }
    }
    stv<C>(p + c, pv);
    stv<C>(r_new + c, rout);
    stv<C>(q + c, qout);
  }
};

}  // namespace

}  // namespace euler
