"""Device arithmetic checked on the CPU.  The row operators of the pressure-solve kernels
(euler_b200/csrc/pcg_ops.cuh) and the masked bilinear sampler of the advection kernels
(csrc/interp.cuh) — the SAME source the GPU kernels instantiate — are compiled for the host and
checked BIT FOR BIT against the oracle, without a GPU: red-black forward / backward solves, the fused search-direction update + A·s,
the unfused A·s and the mixed-precision residual replacement, for both storage types and both
cells-per-thread variants.  What is not covered here is the machinery that feeds the operators
on the device (TMA ring, work split, grid reductions): tests/test_gpu_*.py.

The fused dot products are accumulated by one operator instance in row-major order here, which
is the oracle's sequential dot(): they must match exactly too."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, same_bits
from euler_b200 import shipped_text, resample
from oracle.oracle import Oracle, PRECON_REDBLACK

GUARD = 4
LIB = os.path.join(ROOT, "build", "libkernels_host.so")


@pytest.fixture(scope="module")
def ops():
    subprocess.run(["make", "-C", ROOT, "hostops"], check=True, capture_output=True)
    L = C.CDLL(LIB)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    L.ops_rb_forward_f64.argtypes = L.ops_rb_forward_f32.argtypes = [i, i, i, i, vp, vp, vp, vp]
    L.ops_rb_backward_f64.argtypes = L.ops_rb_backward_f32.argtypes = [i, i, i, i, vp, vp, vp, vp, vp]
    L.ops_rb_backward_f64.restype = L.ops_rb_backward_f32.restype = d
    L.ops_fused_search_apply_f64.argtypes = L.ops_fused_search_apply_f32.argtypes = [i, i, i, i, vp, vp, vp, vp, d, i, vp, vp]
    L.ops_fused_search_apply_f64.restype = L.ops_fused_search_apply_f32.restype = d
    L.ops_apply_a_f64.argtypes = [i, i, i, vp, vp, vp, vp]
    L.ops_apply_a_f64.restype = d
    L.ops_true_residual.argtypes = [i, i, i, vp, vp, vp, vp, vp]
    L.ops_fused_axpy_forward_f64.argtypes = [i, i, i, vp, vp, vp, vp, vp, d, vp, vp, vp]
    L.ops_fused_axpy_forward_f64.restype = d
    return L


class Planes:
    """Device layout on the host: pitch = nx rounded up to 32, GUARD zero rows above and below."""

    def __init__(self, nx, ny):
        self.nx, self.ny, self.pitch = nx, ny, (nx + 31) // 32 * 32
        self.keep = []

    def put(self, a):
        buf = np.zeros((self.ny + 2 * GUARD, self.pitch), dtype=a.dtype)
        buf[GUARD:GUARD + self.ny, :self.nx] = a
        self.keep.append(buf)
        return buf

    def ptr(self, buf):
        return buf.ctypes.data + GUARD * self.pitch * buf.itemsize

    def get(self, buf):
        return buf[GUARD:GUARD + self.ny, :self.nx]


def _state(name, nx, ny, frames):
    """An oracle in red-black mode a few frames in, with rhs and a_diag of the next solve built."""
    text = shipped_text(name) if (nx, ny) == (100, 40) else resample(shipped_text(name), nx - 2, ny - 2)
    o = Oracle(nx, ny, text)
    o.c.precon_mode = PRECON_REDBLACK
    o.c.quirk_marker_dt_leak = 0
    for _ in range(frames):
        o.step_frame()
    dt = o.calculate_timestep(0.1)
    o.substep(dt)
    o.build_rhs(o.calculate_timestep(0.1))
    assert (o.count != 0).sum() > 50 and np.abs(o.b).max() > 0
    return o


CASES = [("block", 100, 40, 14), ("waterfall", 100, 40, 20), ("weird-edges", 100, 40, 15),
         ("filter", 160, 90, 12), ("waterfall", 600, 70, 10)]      # 600 wide: two 512-cell tiles per row


@pytest.mark.parametrize("cpt", [2, 4])
@pytest.mark.parametrize("name,nx,ny,frames", CASES)
def test_fp64_operators_bit_exact(ops, name, nx, ny, frames, cpt):
    o = _state(name, nx, ny, frames)
    fl = o.count != 0
    P = Planes(nx, ny)
    fluid, adiag = P.put(o.count), P.put(o.adiag)
    # z = M^-1 r: forward and backward solves, z.r
    o.r[:] = o.b
    o.apply_preconditioner(o.r, o.z)
    r, pc = P.put(o.r), P.put(np.where(fl, o.precon, 0.0))
    q, z = P.put(np.zeros_like(o.r)), P.put(np.zeros_like(o.r))
    ops.ops_rb_forward_f64(nx, ny, P.pitch, cpt, P.ptr(r), P.ptr(pc), P.ptr(fluid), P.ptr(q))
    assert same_bits(P.get(q)[fl], o.q[fl])
    zr = ops.ops_rb_backward_f64(nx, ny, P.pitch, cpt, P.ptr(q), P.ptr(pc), P.ptr(r), P.ptr(fluid), P.ptr(z))
    assert same_bits(P.get(z)[fl], o.z[fl])
    assert zr == o.dot(o.z, o.r)
    # s' = z + beta s ; A s' ; (A s').s'
    rng = np.random.default_rng(3)
    s_old = rng.standard_normal((ny, nx))
    beta = 0.8125 + 1e-3 * rng.random()
    for init in (1, 0):
        s_ref = np.where(fl, o.z if init else o.z + beta * s_old, s_old)
        as_ref = np.zeros_like(s_ref)
        o.apply_a(np.ascontiguousarray(s_ref), as_ref)
        s_in, s_new, a_s = P.put(s_old), P.put(np.zeros_like(s_old)), P.put(np.zeros_like(s_old))
        acc = ops.ops_fused_search_apply_f64(nx, ny, P.pitch, cpt, P.ptr(z), P.ptr(s_in), P.ptr(fluid), P.ptr(adiag),
                                             beta, init, P.ptr(s_new), P.ptr(a_s))
        assert same_bits(P.get(s_new)[fl], s_ref[fl])
        assert same_bits(P.get(a_s)[fl], as_ref[fl])
        assert acc == o.dot(as_ref, np.ascontiguousarray(s_ref))
    # the unfused operator
    out, out_ref = P.put(np.zeros_like(s_old)), np.zeros_like(s_old)
    o.apply_a(s_old, out_ref)
    acc = ops.ops_apply_a_f64(nx, ny, P.pitch, P.ptr(P.put(s_old)), P.ptr(fluid), P.ptr(adiag), P.ptr(out))
    assert same_bits(P.get(out)[fl], out_ref[fl]) and acc == o.dot(out_ref, s_old)


@pytest.mark.parametrize("cpt", [2, 4])
@pytest.mark.parametrize("name,nx,ny,frames", CASES)
def test_fp32_operators_bit_exact(ops, name, nx, ny, frames, cpt):
    o = _state(name, nx, ny, frames)
    fl = o.count != 0
    P = Planes(nx, ny)
    fluid, adiag = P.put(o.count), P.put(o.adiag)
    o.r32[:] = o.b.astype(np.float32)
    o.rb_build32()
    o.rb_apply32(o.r32, o.z32)
    r, pc = P.put(o.r32), P.put(np.where(fl, o.pc32, np.float32(0)))
    q, z = P.put(np.zeros_like(o.r32)), P.put(np.zeros_like(o.r32))
    ops.ops_rb_forward_f32(nx, ny, P.pitch, cpt, P.ptr(r), P.ptr(pc), P.ptr(fluid), P.ptr(q))
    assert same_bits(P.get(q)[fl], o.q32[fl])
    zr = ops.ops_rb_backward_f32(nx, ny, P.pitch, cpt, P.ptr(q), P.ptr(pc), P.ptr(r), P.ptr(fluid), P.ptr(z))
    assert same_bits(P.get(z)[fl], o.z32[fl])
    # fp64 products of the fp32 values, summed sequentially in row-major order (cumsum is sequential)
    assert zr == float((o.z32[fl].astype(np.float64) * o.r32[fl].astype(np.float64)).cumsum()[-1])
    rng = np.random.default_rng(4)
    s_old = rng.standard_normal((ny, nx)).astype(np.float32)
    beta = 0.8125 + 1e-3 * rng.random()
    for init in (1, 0):
        s_ref = np.where(fl, o.z32 if init else o.z32 + np.float32(beta) * s_old, s_old).astype(np.float32)
        as_ref = np.zeros_like(s_ref)
        o.apply_a32(np.ascontiguousarray(s_ref), as_ref)
        s_in, s_new, a_s = P.put(s_old), P.put(np.zeros_like(s_old)), P.put(np.zeros_like(s_old))
        acc = ops.ops_fused_search_apply_f32(nx, ny, P.pitch, cpt, P.ptr(z), P.ptr(s_in), P.ptr(fluid), P.ptr(adiag),
                                             beta, init, P.ptr(s_new), P.ptr(a_s))
        assert same_bits(P.get(s_new)[fl], s_ref[fl])
        assert same_bits(P.get(a_s)[fl], as_ref[fl])
        seq = (as_ref[fl].astype(np.float64) * s_ref[fl].astype(np.float64)).cumsum()[-1]
        assert acc == float(seq)


def test_true_residual_bit_exact(ops):
    o = _state("waterfall", 100, 40, 25)
    fl = o.count != 0
    P = Planes(100, 40)
    p = np.where(fl, np.random.default_rng(9).standard_normal(o.p.shape) * 300.0, 0.0)
    ap = np.zeros_like(p)
    o.apply_a(p, ap)
    ref = (o.b - ap).astype(np.float32)
    r = P.put(np.zeros((40, 100), np.float32))
    ops.ops_true_residual(100, 40, P.pitch, P.ptr(P.put(p)), P.ptr(P.put(o.b)), P.ptr(P.put(o.count)),
                          P.ptr(P.put(o.adiag)), P.ptr(r))
    assert same_bits(P.get(r)[fl], ref[fl])
    assert not P.get(r)[~fl].any()


def _random_state(nx, ny, seed, density):
    """Arbitrary ragged masks: random fluid (never on the border ring), random solids, a_diag as
    build_rhs would leave it — isolated cells, cells on both sides of the 512-column tile edges,
    rows that end inside the pitch padding."""
    rng = np.random.default_rng(seed)
    o = Oracle(nx, ny, "")
    o.c.precon_mode = PRECON_REDBLACK
    solid = (rng.random((ny, nx)) < 0.08).astype(np.uint8)
    cnt = ((rng.random((ny, nx)) < density) & (solid == 0)).astype(np.uint8) * rng.integers(1, 5, (ny, nx)).astype(np.uint8)
    cnt[0] = cnt[-1] = 0; cnt[:, 0] = cnt[:, -1] = 0
    o.solid[:] = solid; o.count[:] = cnt
    s = solid.astype(np.int32)
    a = np.zeros((ny, nx), np.int32)
    a[1:-1, 1:-1] = 4 - s[1:-1, :-2] - s[1:-1, 2:] - s[:-2, 1:-1] - s[2:, 1:-1]
    o.adiag[:] = np.where(cnt != 0, a, 0).astype(np.int8)
    return o, rng


@pytest.mark.parametrize("nx,ny,density,seed", [(1030, 37, 0.55, 1), (513, 20, 0.15, 2), (96, 70, 0.9, 3), (544, 9, 0.5, 4)])
def test_operators_on_random_ragged_masks(ops, nx, ny, density, seed):
    o, rng = _random_state(nx, ny, seed, density)
    fl = o.count != 0
    P = Planes(nx, ny)
    fluid, adiag = P.put(o.count), P.put(o.adiag)
    scale = 10.0 ** rng.integers(-3, 4, (ny, nx))
    # fp64
    o.r[:] = np.where(fl, rng.standard_normal((ny, nx)) * scale, 0.0)
    o.apply_preconditioner(o.r, o.z)
    r, pc = P.put(o.r), P.put(np.where(fl, o.precon, 0.0))
    q, z = P.put(np.zeros((ny, nx))), P.put(np.zeros((ny, nx)))
    for cpt in (2, 4):
        ops.ops_rb_forward_f64(nx, ny, P.pitch, cpt, P.ptr(r), P.ptr(pc), P.ptr(fluid), P.ptr(q))
        zr = ops.ops_rb_backward_f64(nx, ny, P.pitch, cpt, P.ptr(q), P.ptr(pc), P.ptr(r), P.ptr(fluid), P.ptr(z))
        assert same_bits(P.get(q)[fl], o.q[fl]) and same_bits(P.get(z)[fl], o.z[fl]) and zr == o.dot(o.z, o.r)
    # fp32
    o.r32[:] = o.r.astype(np.float32)
    o.rb_build32(); o.rb_apply32(o.r32, o.z32)
    r, pc = P.put(o.r32), P.put(np.where(fl, o.pc32, np.float32(0)))
    q, z = P.put(np.zeros((ny, nx), np.float32)), P.put(np.zeros((ny, nx), np.float32))
    for cpt in (2, 4):
        ops.ops_rb_forward_f32(nx, ny, P.pitch, cpt, P.ptr(r), P.ptr(pc), P.ptr(fluid), P.ptr(q))
        ops.ops_rb_backward_f32(nx, ny, P.pitch, cpt, P.ptr(q), P.ptr(pc), P.ptr(r), P.ptr(fluid), P.ptr(z))
        assert same_bits(P.get(q)[fl], o.q32[fl]) and same_bits(P.get(z)[fl], o.z32[fl])
    # fused search + apply, both types
    for dt, fn, apply in ((np.float64, ops.ops_fused_search_apply_f64, o.apply_a), (np.float32, ops.ops_fused_search_apply_f32, o.apply_a32)):
        zz = (rng.standard_normal((ny, nx)) * scale).astype(dt)
        ss = (rng.standard_normal((ny, nx)) * scale).astype(dt)
        beta = 1.37
        s_ref = np.where(fl, zz + dt(beta) * ss, ss).astype(dt)
        as_ref = np.zeros((ny, nx), dt)
        apply(np.ascontiguousarray(s_ref), as_ref)
        s_new, a_s = P.put(np.zeros((ny, nx), dt)), P.put(np.zeros((ny, nx), dt))
        fn(nx, ny, P.pitch, 4, P.ptr(P.put(zz)), P.ptr(P.put(ss)), P.ptr(fluid), P.ptr(adiag), beta, 0, P.ptr(s_new), P.ptr(a_s))
        assert same_bits(P.get(s_new)[fl], s_ref[fl]) and same_bits(P.get(a_s)[fl], as_ref[fl])


def test_fused_axpy_forward_variant_bit_exact(ops):
    """stencil_variant 2: p, r', ||r'||inf and q = L^-1 r' in one pass equal the separate
    fmadd (main.c:753-754), inf_norm (:756) and forward solve."""
    o = _state("waterfall", 160, 90, 18)
    nx, ny = 160, 90
    fl = o.count != 0
    rng = np.random.default_rng(12)
    o.r[:] = o.b
    o.apply_preconditioner(o.r, o.z)                       # builds o.precon
    s = np.where(fl, rng.standard_normal((ny, nx)), 0.0)
    a_s = np.zeros((ny, nx)); o.apply_a(s, a_s)
    p0 = np.where(fl, rng.standard_normal((ny, nx)) * 50.0, 0.0)
    alpha = 0.3731
    p_ref = np.where(fl, p0 + s * alpha, p0)
    r_ref = np.where(fl, o.b + a_s * -alpha, o.b)
    zz = np.zeros((ny, nx)); o.apply_preconditioner(np.ascontiguousarray(r_ref), zz)    # o.q = L^-1 r'
    P = Planes(nx, ny)
    p, r_new, q = P.put(p0), P.put(np.zeros((ny, nx))), P.put(np.zeros((ny, nx)))
    mx = ops.ops_fused_axpy_forward_f64(nx, ny, P.pitch, P.ptr(P.put(o.b)), P.ptr(P.put(a_s)),
                                        P.ptr(P.put(np.where(fl, o.precon, 0.0))), P.ptr(P.put(o.count)),
                                        P.ptr(P.put(s)), alpha, P.ptr(p), P.ptr(r_new), P.ptr(q))
    assert same_bits(P.get(p)[fl], p_ref[fl]) and same_bits(P.get(r_new)[fl], r_ref[fl])
    assert same_bits(P.get(q)[fl], o.q[fl])
    assert mx == float(np.abs(r_ref[fl]).max())


@pytest.mark.parametrize("nx,ny,seed", [(100, 40, 1), (37, 53, 2), (257, 19, 3)])
def test_masked_bilinear_sampler_bit_exact(ops, nx, ny, seed):
    """interp.cuh's interpolate<P|U|V> (the sampler of advect_u/v, advect_markers, advect_p)
    against the oracle's restatement of main.c:301-364 on random planes, ragged fluid masks and
    sample positions that include exact grid lines, the clamped range ends and far outliers."""
    ops.ops_interpolate.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(seed)
    o = Oracle(nx, ny, "")
    cnt = (rng.random((ny, nx)) < 0.6).astype(np.uint8) * rng.integers(1, 5, (ny, nx)).astype(np.uint8)
    cnt[0] = cnt[-1] = 0; cnt[:, 0] = cnt[:, -1] = 0
    o.count[:] = cnt
    q = rng.uniform(-30, 30, (ny, nx)).astype(np.float32)
    n = 4000
    ix = rng.uniform(-3, nx + 3, n).astype(np.float32)
    iy = rng.uniform(-3, ny + 3, n).astype(np.float32)
    ix[:400] = np.floor(ix[:400]); iy[200:600] = np.floor(iy[200:600])          # on grid lines
    special = np.array([0, -0.0, nx - 1, nx - 2, ny - 1, ny - 2, 1e9, -1e9, 0.5, nx - 1.5], np.float32)
    ix[600:610] = special; iy[605:615] = special
    P = Planes(nx, ny)
    qp, fp = P.put(q), P.put(cnt)
    for t in (0, 1, 2):
        out = np.empty(n, np.float32)
        ops.ops_interpolate(nx, ny, P.pitch, t, P.ptr(qp), P.ptr(fp), n, ix.ctypes.data, iy.ctypes.data, out.ctypes.data)
        ref = np.array([o.interpolate(q, float(a), float(b), t) for a, b in zip(ix, iy)], np.float32)
        assert same_bits(out, ref), "type %d: %d samples differ" % (t, int((out.view(np.uint32) != ref.view(np.uint32)).sum()))


def test_rng_jump_ahead_equals_sequential_draws(ops):
    """csrc/rng.cuh: jumping k draws ahead (what gives every source cell its own slice of the
    stream) lands on the state k sequential xorshift64 steps reach, and the floats are the host
    program's randf() (euler_b200/host/scenario.c == reference main.c:203-207, misc/rng.c)."""
    from euler_b200 import scenario as S
    ops.ops_rng_jump.restype = C.c_uint64
    ops.ops_rng_jump.argtypes = [C.c_uint64, C.c_uint64]
    ops.ops_rng_draws.restype = C.c_uint64
    ops.ops_rng_draws.argtypes = [C.c_uint64, C.c_int, C.c_void_p]
    seed = 0x9bd185c449534b91
    n = 5000
    out = np.empty(n, np.float32)
    end = ops.ops_rng_draws(seed, n, out.ctypes.data)
    host_state = C.c_uint64(seed)
    host = np.array([S._lib().euler_randf(C.byref(host_state)) for _ in range(n)], np.float32)
    assert same_bits(out, host) and host_state.value == end
    assert ops.ops_rng_jump(seed, 0) == seed and ops.ops_rng_jump(seed, n) == end
    # composition and large strides: jump(a + b) == jump(jump(a), b); 2^40 draws in one go
    rng = np.random.default_rng(5)
    for _ in range(50):
        a, b = int(rng.integers(0, 2 ** 40)), int(rng.integers(0, 2 ** 40))
        s = int(rng.integers(1, 2 ** 63))
        assert ops.ops_rng_jump(ops.ops_rng_jump(s, a), b) == ops.ops_rng_jump(s, a + b)
    k = 123457
    tmp = np.empty(k, np.float32)
    assert ops.ops_rng_jump(seed, k) == ops.ops_rng_draws(seed, k, tmp.ctypes.data)


@pytest.mark.parametrize("name,nx,ny,frames,dt", [("block", 100, 40, 16, None), ("waterfall", 100, 40, 30, None),
                                                   ("weird-edges", 100, 40, 20, 0.3), ("filter", 160, 90, 15, 0.25)])
def test_marker_walk_bit_exact(ops, name, nx, ny, frames, dt):
    """csrc/marker_walk.cuh: the RK1 step with the reference's grid-line walk and solid rewind
    (main.c:464-537) and the marker -> cell map (main.c:106-107), marker by marker, against the
    oracle (per-marker dt): positions bit-exact, cells equal.  With `dt` given the markers are
    re-sprinkled over the whole domain and pushed with an over-long step through random
    velocities, so that walks cross many cells and run into solids."""
    ops.ops_walk_markers.argtypes = [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_float, C.c_float, C.c_int] + [C.c_void_p] * 3
    text = shipped_text(name) if (nx, ny) == (100, 40) else resample(shipped_text(name), nx - 2, ny - 2)
    o = Oracle(nx, ny, text)
    o.c.quirk_marker_dt_leak = 0
    for _ in range(frames):
        o.step_frame()
    rng = np.random.default_rng(8)
    if dt is None:
        dt = o.calculate_timestep(0.1)
    else:
        k = 12000                                            # < MAX_MARKER_COUNT = 4 nx ny
        m = np.stack([rng.uniform(1.01, nx - 1.01, k), rng.uniform(1.01, ny - 1.01, k)], 1).astype(np.float32)
        o.set_markers(m)
        o.u[:] = rng.uniform(-3, 3, (ny, nx)).astype(np.float32)
        o.v[:] = rng.uniform(-3, 3, (ny, nx)).astype(np.float32)
    src = o.markers.copy()
    n = len(src)
    assert n > 300
    P = Planes(nx, ny)
    u, v, fluid, solid = P.put(o.u), P.put(o.v), P.put(o.count), P.put(o.solid)
    dst = np.empty_like(src)
    cells = np.empty(n, np.int64)
    ops.ops_walk_markers(nx, ny, P.pitch, P.ptr(u), P.ptr(v), P.ptr(fluid), P.ptr(solid), 1.0, dt, n,
                         src.ctypes.data, dst.ctypes.data, cells.ctypes.data)
    o.advect_markers(dt)
    ref = o.markers
    ok = np.isfinite(ref).all(1)
    assert ok.sum() > 0.9 * n
    assert same_bits(dst[ok], ref[ok])
    ref_cells = np.floor(ref[ok, 1]).astype(np.int64) * nx + np.floor(ref[ok, 0]).astype(np.int64)
    inside = (ref[ok, 0] >= 0) & (ref[ok, 0] < nx) & (ref[ok, 1] >= 0) & (ref[ok, 1] < ny)
    assert np.array_equal(cells[ok][inside], ref_cells[inside])


def _nan_equal_bits(a, b):
    """bit-exact except that NaNs (0/0 means of extrapolate, assert off) only have to coincide"""
    return np.array_equal(np.isnan(a), np.isnan(b)) and same_bits(np.nan_to_num(a), np.nan_to_num(b))


GRID_CASES = [("block", 100, 40, 16), ("waterfall", 100, 40, 30), ("weird-edges", 100, 40, 20),
              ("filter", 160, 90, 15), ("waterfall", 333, 129, 12)]


@pytest.mark.parametrize("name,nx,ny,frames", GRID_CASES)
def test_grid_stage_quads_bit_exact(ops, name, nx, ny, frames):
    """csrc/grid_ops.cuh (the per-quad arithmetic of k_extrapolate_bounds, k_advect_velocity,
    k_build_rhs, k_pressure_update) in sim_step order from an oracle state: every output plane
    bit-exact against the oracle's stages (main.c:865-889, 713-733, 769-805)."""
    vp, i, f, d = C.c_void_p, C.c_int, C.c_float, C.c_double
    ops.ops_extrapolate_bounds.argtypes = [i, i, i] + [vp] * 7
    ops.ops_advect_velocity.argtypes = [i, i, i] + [vp] * 4 + [f, f, f, vp, vp]
    ops.ops_build_rhs.argtypes = [i, i, i] + [vp] * 4 + [f, d, vp, vp]
    ops.ops_build_rhs.restype = i
    ops.ops_pressure_update.argtypes = [i, i, i] + [vp] * 5 + [f, f, vp, vp]
    text = shipped_text(name) if (nx, ny) == (100, 40) else resample(shipped_text(name), nx - 2, ny - 2)
    o = Oracle(nx, ny, text)
    o.c.precon_mode = PRECON_REDBLACK
    o.c.quirk_marker_dt_leak = 0
    for _ in range(frames):
        o.step_frame()
    # the marker half of a sub-step on the oracle, then the grid stages on both sides
    dt = o.calculate_timestep(0.1)
    o.advect_markers(dt); o.refresh_marker_counts(); o.update_fluid_sources()
    P = Planes(nx, ny)
    fluid, prev, solid = P.put(o.count), P.put(o.prev_count), P.put(o.solid)
    u, v = P.put(o.u), P.put(o.v)
    ue, ve = P.put(np.zeros_like(o.u)), P.put(np.zeros_like(o.u))
    ops.ops_extrapolate_bounds(nx, ny, P.pitch, P.ptr(u), P.ptr(v), P.ptr(fluid), P.ptr(prev), P.ptr(solid), P.ptr(ue), P.ptr(ve))
    o.extrapolate(o.u, 1); o.extrapolate(o.v, 2); o.zero_bounds(o.u, 1); o.zero_bounds(o.v, 2)
    assert _nan_equal_bits(P.get(ue), o.u) and _nan_equal_bits(P.get(ve), o.v)
    o.u[:] = np.nan_to_num(o.u); o.v[:] = np.nan_to_num(o.v)
    u, v = P.put(o.u), P.put(o.v)
    ut, vt = P.put(np.zeros_like(o.u)), P.put(np.zeros_like(o.u))
    ops.ops_advect_velocity(nx, ny, P.pitch, P.ptr(u), P.ptr(v), P.ptr(fluid), P.ptr(solid), dt, 1.0, -10.0, P.ptr(ut), P.ptr(vt))
    o.advect_u(dt); o.advect_v(dt); o.apply_body_forces(dt); o.zero_bounds(o.utmp, 1); o.zero_bounds(o.vtmp, 2)
    assert same_bits(P.get(ut), o.utmp) and same_bits(P.get(vt), o.vtmp)
    # rhs + a_diag
    fl = o.count != 0
    scale = float(np.float32(np.float32(1.0) * np.float32(1.0) * np.float32(1.0) / np.float32(dt)))   # main.c:713, fp32
    b, adiag = P.put(np.zeros((ny, nx))), P.put(o.adiag.copy())
    nonzero = ops.ops_build_rhs(nx, ny, P.pitch, P.ptr(ut), P.ptr(vt), P.ptr(fluid), P.ptr(solid), 1.0, scale, P.ptr(b), P.ptr(adiag))
    o.build_rhs(dt)
    assert same_bits(P.get(b), o.b) and np.array_equal(P.get(adiag)[fl], o.adiag[fl])
    assert bool(nonzero) == bool(np.any(o.b[fl] != 0))
    # pressure update from the oracle's own solve
    o.project(dt)                                     # leaves the CLAMPED p; redo with an unclamped one
    rng = np.random.default_rng(2)
    p_raw = np.where(fl, o.p - rng.random((ny, nx)) * 0.3 * (rng.random((ny, nx)) < 0.2), 0.0)   # some negatives
    o.p[:] = p_raw
    o.pressure_update(dt)
    p = P.put(p_raw)
    uo, vo = P.put(np.zeros_like(o.u)), P.put(np.zeros_like(o.u))
    ops.ops_pressure_update(nx, ny, P.pitch, P.ptr(p), P.ptr(ut), P.ptr(vt), P.ptr(fluid), P.ptr(solid), dt, 1.0, P.ptr(uo), P.ptr(vo))
    assert same_bits(P.get(uo), o.u) and same_bits(P.get(vo), o.v)
    assert same_bits(P.get(p)[fl], o.p[fl]) and float(P.get(p)[fl].min()) >= 0.0
