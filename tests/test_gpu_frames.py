"""GPU parity over whole frames (euler_gpu_step_frame == sim_step, main.c:843-900) and
size-independent properties at BASELINE.json's larger sizes."""
import numpy as np
import pytest

from conftest import SCENARIOS, same_bits
from euler_b200 import Scenario, shipped_text, resample, synthetic

pytestmark = pytest.mark.gpu


def _pair(text, nx, ny, precon, leak, dot_mode=0, **kw):
    from euler_b200 import gpu as G
    from oracle.oracle import Oracle
    o = Oracle(nx, ny, text)
    o.c.precon_mode = precon
    o.c.quirk_marker_dt_leak = leak
    g = G.EulerGpu.from_scenario(Scenario(text, nx, ny), precon=precon, dot_mode=dot_mode,
                                 marker_mode=G.MARKERS_REFERENCE if leak else G.MARKERS_FAST, **kw)
    return o, g, G


@pytest.mark.parametrize("name", SCENARIOS)
def test_frames_ic0_wavefront_bit_exact_classification(name):
    """Reference-faithful mode: the count plane (cell classification) and the marker array stay
    bit-exact; velocities stay within 1e-5 relative (they are in fact usually bit-exact: the
    only non-bit-exact ingredient is the summation order of the PCG dot products)."""
    o, g, G = _pair(shipped_text(name), 100, 40, 0, 0)
    for f in range(25):
        so, sg = o.step_frame(), g.step_frame()
        assert so == sg
        assert same_bits(g.get(G.F_COUNT), o.count), "frame %d" % f
    assert same_bits(g.get(G.F_MARKERS), o.markers)
    for fld, ref in ((G.F_U, o.u), (G.F_V, o.v)):
        assert float(np.abs(g.get(fld) - ref).max()) <= 1e-5 * max(1.0, float(np.abs(ref).max()))
    g.close()


@pytest.mark.parametrize("name", SCENARIOS)
def test_frames_fully_bit_exact_with_reference_order_dots(name):
    """IC(0) wavefront + reference-order dot products: EVERYTHING is bit-identical to the CPU
    over 40 frames — counts, marker array, u, v, precon, iteration totals."""
    o, g, G = _pair(shipped_text(name), 100, 40, 0, 0, dot_mode=1)
    for f in range(40):
        assert o.step_frame() == g.step_frame()
    assert same_bits(g.get(G.F_COUNT), o.count) and same_bits(g.get(G.F_MARKERS), o.markers)
    assert same_bits(g.get(G.F_U), o.u) and same_bits(g.get(G.F_V), o.v)
    assert same_bits(g.get(G.F_PRECON), o.precon)
    st = g.stats()
    assert st.pcg_iterations == o.c.total_iterations and st.solves == o.c.total_solves
    g.close()


@pytest.mark.parametrize("name", ["block", "waterfall"])
def test_frames_red_black(name):
    """Red-black mode against its CPU mirror: same iteration counts early on, classification
    equal, velocities within 1e-5 over the first frames."""
    o, g, G = _pair(shipped_text(name), 100, 40, 1, 0)
    for f in range(8):
        o.step_frame(); g.step_frame()
    assert same_bits(g.get(G.F_COUNT), o.count)
    for fld, ref in ((G.F_U, o.u), (G.F_V, o.v)):
        assert float(np.abs(g.get(fld) - ref).max()) <= 1e-5 * max(1.0, float(np.abs(ref).max()))
    g.close()


def test_red_black_converges_to_same_tolerance_as_ic0():
    """Both preconditioners, iteration cap raised, reach ||r||inf <= 1e-6 and the same p."""
    from euler_b200 import gpu as G
    text = shipped_text("block")
    scn = Scenario(text, 100, 40)
    sims = [G.EulerGpu.from_scenario(scn, precon=p, marker_mode=G.MARKERS_FAST, max_iterations=2000)
            for p in (0, 1)]
    for s in sims:
        for _ in range(14):
            s.step_frame()
    u = sims[0].get(G.F_UTMP); v = sims[0].get(G.F_VTMP); c = sims[0].get(G.F_COUNT)
    sims[1].set(G.F_UTMP, u); sims[1].set(G.F_VTMP, v); sims[1].set(G.F_COUNT, c)
    ps = []
    for s in sims:
        s.run_stage(G.S_PROJECT, 0.02)
        st = s.stats()
        assert 0 < st.last_iterations < 2000 and st.last_residual <= 1e-6
        ps.append(s.get(G.F_P))
    fl = c != 0
    assert float(np.abs(ps[0][fl] - ps[1][fl]).max()) <= 1e-5 * max(1.0, float(np.abs(ps[0][fl]).max()))
    for s in sims:
        s.close()


def test_1024_ic0_wavefront_one_substep_vs_oracle():
    """BASELINE config[1]: block scenario upscaled to 1024^2, IC(0) wavefront mode.  The
    reference hits its 100-iteration cap here (||r||inf ~ 10 at exit): unconverged CG amplifies
    the 1e-16 differences of a tree-reduced dot product to ~5e-6 in p (and |p|*1e-5 in u, v), so
    this configuration runs with reference-order dot products and must be BIT-exact."""
    n = 1024
    text = synthetic("basic-fill", n, n)
    o, g, G = _pair(text, n, n, 0, 0, dot_mode=1)
    dt = o.calculate_timestep(0.1)
    assert g.calculate_timestep(0.1) == dt
    o.substep(dt); g.substep(dt)
    st = g.stats()
    assert st.last_iterations == o.c.last_iterations == 100
    assert same_bits(g.get(G.F_COUNT), o.count) and same_bits(g.get(G.F_MARKERS), o.markers)
    fl = o.count != 0
    p = g.get(G.F_P)
    assert same_bits(p[fl], o.p[fl])
    assert same_bits(g.get(G.F_U), o.u) and same_bits(g.get(G.F_V), o.v)
    g.close()
    # the same sub-step with tree-reduced dots stays within the stated 1e-5 on pressure
    o2, g2, G = _pair(text, n, n, 0, 0, dot_mode=0)
    o2.substep(dt); g2.substep(dt)
    p2 = g2.get(G.F_P)
    assert float(np.abs(p2[fl] - o2.p[fl]).max()) <= 1e-5 * float(np.abs(o2.p[fl]).max())
    g2.close()


def test_4096_properties():
    """Size-independent properties at 4096^2 (no oracle run at this size): the count plane sums
    to the marker count, PCG reduces the residual, IC(0) and red-black agree on A*s, pressure is
    non-negative after the clamp, solid faces carry no flow."""
    from euler_b200 import gpu as G
    n = 4096
    text = resample(shipped_text("waterfall"), n - 2, n - 2)
    scn = Scenario(text, n, n)
    g = G.EulerGpu.from_scenario(scn, precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST)
    for _ in range(2):
        g.step_frame()
    st = g.stats()
    cnt = g.get(G.F_COUNT)
    assert int(cnt.astype(np.int64).sum()) == int(st.n_markers)
    solid = scn.solid != 0
    u, v = g.get(G.F_U), g.get(G.F_V)
    assert not u[:, :-1][solid[:, :-1] | solid[:, 1:]].any()
    assert not v[:-1][solid[:-1] | solid[1:]].any()
    p = g.get(G.F_P)
    assert float(p[cnt != 0].min()) >= 0.0
    assert np.isfinite(u).all() and np.isfinite(v).all()
    g.close()


def test_host_program_headless_matches_golden(known_answers, tmp_path):
    """bin/euler-gpu (the C host: parser, seeding, loop) run headless on the shipped scenarios
    reproduces the reference's known-answer hashes of the count plane."""
    import json
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "bin", "euler-gpu")
    assert os.path.exists(exe), "run `make host`"
    for name, frames in (("block", 10), ("block", 50), ("basic", 50), ("weird-edges", 50), ("waterfall", 10)):
        path = tmp_path / (name + ".txt")
        path.write_bytes(shipped_text(name))
        out = subprocess.run([exe, "--headless", "--frames", str(frames), "--exact-dot", str(path)],
                             check=True, capture_output=True, text=True, cwd=ROOT).stdout
        res = json.loads(out.strip().splitlines()[-1])
        want = known_answers["%s@100x40/f%d" % (name, frames)]
        assert res["fnv_count"] == want["fnv_count"], (name, frames)
        assert res["markers"] == want["markers"] and res["rng_state"] == want["rng_state"]


@pytest.mark.parametrize("name,frames", [("filter", 60), ("waterfall", 80), ("block", 30)])
def test_reference_marker_mode_reproduces_dt_carry_over(name, frames):
    """EULER_MARKERS_REFERENCE: the `dt -= t_prev` carry-over between successive markers
    (main.c:464,501,518) is reproduced — with reference-order dots the run stays bit-identical
    to the oracle (quirk ON) well past the frame where the quirk first fires (filter: 29,
    waterfall: 46)."""
    o, g, G = _pair(shipped_text(name), 100, 40, 0, 1, dot_mode=1)
    for f in range(frames):
        assert o.step_frame() == g.step_frame()
        assert same_bits(g.get(G.F_COUNT), o.count), "frame %d" % (f + 1)
    assert same_bits(g.get(G.F_MARKERS), o.markers)
    assert same_bits(g.get(G.F_U), o.u) and same_bits(g.get(G.F_V), o.v)
    g.close()


def test_known_answers_through_gpu(known_answers):
    """The reference's own known answers (tests/golden/known_answers.json, frame 50) straight
    from the GPU in its default-faithful configuration."""
    from oracle.oracle import fnv1a
    from euler_b200 import gpu as G
    for name in SCENARIOS:
        g = G.EulerGpu.from_scenario(Scenario(shipped_text(name), 100, 40), dot_mode=1)
        for _ in range(50):
            g.step_frame()
        want = known_answers["%s@100x40/f50" % name]
        st = g.stats()
        assert "%016x" % fnv1a(g.get(G.F_COUNT)) == want["fnv_count"], name
        assert int(st.n_markers) == want["markers"] and "%016x" % st.rng_state == want["rng_state"]
        assert abs(float(np.abs(g.get(G.F_U).astype(np.float64)).sum()) - want["sum_abs_u"]) < 1e-9
        g.close()


@pytest.mark.parametrize("precon,marker_mode", [(0, 0), (1, 1)])
def test_reinit_equals_fresh_handle(precon, marker_mode):
    """euler_gpu_reinit == sim_init (main.c:209-274) into an existing handle: after dirtying every
    plane (incl. the persistent g_precon) the run that follows is bit-identical to a
    fresh handle's."""
    from euler_b200 import gpu as G
    nx, ny = 160, 90
    text = resample(shipped_text("waterfall"), nx - 2, ny - 2)
    scn = Scenario(text, nx, ny)
    other = Scenario(resample(shipped_text("block"), nx - 2, ny - 2), nx, ny)
    a = G.EulerGpu.from_scenario(other, precon=precon, marker_mode=marker_mode)
    for _ in range(8):
        a.step_frame()
    a.reinit(scn.solid, scn.source, scn.sink, scn.markers, scn.rng_state)
    b = G.EulerGpu.from_scenario(scn, precon=precon, marker_mode=marker_mode)
    for _ in range(20):
        assert a.step_frame() == b.step_frame()
    for f in (G.F_COUNT, G.F_PREV_COUNT, G.F_U, G.F_V, G.F_MARKERS, G.F_PRECON, G.F_P):
        assert same_bits(a.get(f), b.get(f)), f
    assert int(a.stats().rng_state) == int(b.stats().rng_state)
    a.close(); b.close()


# ---- --rainbow colour transport and the renderer's window feed (8f.2) -----------

@pytest.mark.parametrize("name", ["block", "waterfall", "weird-edges", "filter"])
def test_rainbow_frames_bit_exact(name):
    """With reference-order dots the whole run is bit-identical to the oracle (itself pinned to
    the reference's --rainbow run), so the colour planes must be too — whole planes, incl. the
    stale values the reference's plane copies leave at non-fluid cells (main.c:875)."""
    text = shipped_text(name)
    from euler_b200 import gpu as G
    from oracle.oracle import Oracle
    o = Oracle(100, 40, rainbow=True); o.init_from_text(text)
    g = G.EulerGpu.from_scenario(Scenario(text, 100, 40), precon=0, dot_mode=1,
                                 marker_mode=G.MARKERS_REFERENCE, rainbow=1)
    for f in range(25):
        assert g.step_frame() == o.step_frame()
        for fld, plane in ((G.F_CR, o.cr), (G.F_CG, o.cg), (G.F_CB, o.cb), (G.F_COUNT, o.count)):
            assert same_bits(g.get(fld), plane), (name, f, fld)
    o.colorize(); g.colorize()                   # the `r` key, main.c:971-974
    for fld, plane in ((G.F_CR, o.cr), (G.F_CG, o.cg), (G.F_CB, o.cb)):
        assert same_bits(g.get(fld), plane)
    g.close()


def test_rainbow_stages_from_identical_state():
    """extrapolate(P) and advect_p one at a time from the oracle's state, on a resampled grid
    whose pitch is not the row length."""
    from euler_b200 import gpu as G
    from oracle.oracle import Oracle, P
    nx, ny = 333, 129
    text = resample(shipped_text("waterfall"), nx - 2, ny - 2)
    o = Oracle(nx, ny, rainbow=True); o.init_from_text(text)
    o.c.precon_mode = 1; o.c.quirk_marker_dt_leak = 0
    for _ in range(12):
        o.step_frame()
    g = G.EulerGpu.from_scenario(Scenario(text, nx, ny), precon=1, marker_mode=G.MARKERS_FAST, rainbow=1)
    def push():
        g.set(G.F_U, o.u); g.set(G.F_V, o.v); g.set(G.F_COUNT, o.count); g.set(G.F_PREV_COUNT, o.prev_count)
        g.set(G.F_CR, o.cr); g.set(G.F_CG, o.cg); g.set(G.F_CB, o.cb)
    dt = o.calculate_timestep(0.1)
    o.advect_markers(dt); o.refresh_marker_counts()
    push()
    for q in (o.cr, o.cg, o.cb):
        o.extrapolate(q, P)
    g.run_stage(G.S_EXTRAPOLATE_COLOR)
    for fld, plane in ((G.F_CR, o.cr), (G.F_CG, o.cg), (G.F_CB, o.cb)):
        assert same_bits(g.get(fld), plane), fld
    o.update_fluid_sources()
    o.extrapolate(o.u, 1); o.extrapolate(o.v, 2); o.zero_bounds(o.u, 1); o.zero_bounds(o.v, 2)
    push()
    fl = o.count != 0
    for q, t in ((o.cr, o.crtmp), (o.cg, o.cgtmp), (o.cb, o.cbtmp)):
        o.advect_p(q, dt, t)
    g.run_stage(G.S_ADVECT_COLOR, dt)
    for fld, plane in ((G.F_CR, o.crtmp), (G.F_CG, o.cgtmp), (G.F_CB, o.cbtmp)):
        assert same_bits(g.get(fld)[fl], plane[fl]), fld
    g.close()


def test_rainbow_off_and_errors():
    from euler_b200 import gpu as G
    g = G.EulerGpu.from_scenario(Scenario(shipped_text("block"), 100, 40))
    with pytest.raises(G.EulerGpuError):
        g.colorize()
    with pytest.raises(G.EulerGpuError):
        g.get(G.F_CR)
    with pytest.raises(G.EulerGpuError):
        g.run_stage(G.S_ADVECT_COLOR, 0.01)
    g.close()
    # (slab handles carry the colour planes since round 2: tests/test_gpu_multi.py::test_rainbow_on_slabs)


def test_host_program_rainbow_matches_golden(tmp_path):
    """bin/euler-gpu --rainbow --headless: colour-plane hashes equal the reference's known answers."""
    import json, os, subprocess
    from conftest import ROOT, GOLDEN
    with open(os.path.join(GOLDEN, "rainbow_answers.json")) as f:
        want = json.load(f)
    exe = os.path.join(ROOT, "bin", "euler-gpu")
    assert os.path.exists(exe), "run `make host`"
    for name, frames in (("block", 10), ("waterfall", 30), ("weird-edges", 30)):
        path = tmp_path / (name + ".txt")
        path.write_bytes(shipped_text(name))
        out = subprocess.run([exe, "--rainbow", "--headless", "--frames", str(frames), "--exact-dot", str(path)],
                             check=True, capture_output=True, text=True, cwd=ROOT).stdout
        res = json.loads(out.strip().splitlines()[-1])
        w = want["%s@100x40/f%d" % (name, frames)]
        assert res["fnv_count"] == w["fnv_count"], (name, frames)
        assert (res["fnv_r"], res["fnv_g"], res["fnv_b"]) == (w["fnv_r"], w["fnv_g"], w["fnv_b"]), (name, frames)


def test_window_feed_equals_full_plane():
    """euler_gpu_read_window (what draw_rows needs, main.c:917-920) against the full plane."""
    from euler_b200 import gpu as G
    nx, ny = 333, 129
    text = resample(shipped_text("block"), nx - 2, ny - 2)
    g = G.EulerGpu.from_scenario(Scenario(text, nx, ny), precon=1, marker_mode=1)
    for _ in range(6):
        g.step_frame()
    full = g.get(G.F_COUNT)
    fu = g.get(G.F_U)
    for (x0, y0, w, h) in ((1, 69, 80, 59), (0, 0, nx, ny), (17, 5, 1, 1), (300, 100, 33, 29), (5, 5, 0, 0)):
        out = np.full((ny, nx), 255, np.uint8)
        g.read_window(G.F_COUNT, x0, y0, w, h, out)
        assert np.array_equal(out[y0:y0 + h, x0:x0 + w], full[y0:y0 + h, x0:x0 + w])
        mask = np.ones((ny, nx), bool); mask[y0:y0 + h, x0:x0 + w] = False
        assert (out[mask] == 255).all(), "outside the window must stay untouched"
        outf = np.zeros((ny, nx), np.float32)
        g.read_window(G.F_U, x0, y0, w, h, outf)
        assert same_bits(outf[y0:y0 + h, x0:x0 + w], fu[y0:y0 + h, x0:x0 + w])
    with pytest.raises(G.EulerGpuError):
        g.read_window(G.F_COUNT, 300, 0, 40, 10, np.zeros((ny, nx), np.uint8))
    g.close()


@pytest.mark.parametrize("rainbow", [False, True])
def test_host_program_checkpoint_round_trip(rainbow, tmp_path):
    """bin/euler-gpu --save / --load (host/checkpoint.c over euler_gpu_get/set): 12 frames, save,
    load into a fresh process, 13 more frames == 25 frames in one go, bit for bit (count-plane
    hash, RNG state, marker count, colour planes)."""
    import json, os, subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "bin", "euler-gpu")
    path = tmp_path / "waterfall.txt"
    path.write_bytes(shipped_text("waterfall"))
    ck = str(tmp_path / "state.ck")
    base = [exe, "--headless", "--exact-dot"] + (["--rainbow"] if rainbow else [])

    def run(extra):
        out = subprocess.run(base + extra + [str(path)], check=True, capture_output=True, text=True, cwd=ROOT).stdout
        return json.loads(out.strip().splitlines()[-1])
    whole = run(["--frames", "25"])
    run(["--frames", "12", "--save", ck])
    rest = run(["--frames", "13", "--load", ck])
    keys = ["fnv_count", "rng_state", "markers"] + (["fnv_r", "fnv_g", "fnv_b"] if rainbow else [])
    assert [rest[k] for k in keys] == [whole[k] for k in keys]
    # a checkpoint of another grid size / colour mode is refused
    bad = subprocess.run([exe, "--headless", "--frames", "1", "--grid", "64x48", "--load", ck, str(path)],
                         capture_output=True, text=True, cwd=ROOT)
    assert bad.returncode != 0 and "checkpoint" in bad.stderr


def test_4096_operator_is_linear_and_symmetric():
    """Size-independent properties of the stencil kernels at 4096^2 (BASELINE config[2] geometry,
    no oracle run at this size): A is linear and symmetric on the fluid cells, and the red-black
    preconditioner is symmetric too (M^-1 = (L L^T)^-1) — any masking or halo slip in the TMA
    pipeline or the work split would break one of them."""
    from euler_b200 import gpu as G
    n = 4096
    text = resample(shipped_text("waterfall"), n - 2, n - 2)
    g = G.EulerGpu.from_scenario(Scenario(text, n, n), precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST)
    g.step_frame()
    g.run_stage(G.S_BUILD_RHS, 0.01)
    fl = g.get(G.F_COUNT) != 0
    rng = np.random.default_rng(7)
    a = np.where(fl, rng.standard_normal((n, n)), 0.0)
    b = np.where(fl, rng.standard_normal((n, n)), 0.0)

    def apply_a(x):
        g.set(G.F_S, x); g.run_stage(G.S_APPLY_A)
        return np.where(fl, g.get(G.F_Z), 0.0)

    def apply_m(x):
        g.set(G.F_R, x); g.run_stage(G.S_PRECONDITION)
        return np.where(fl, g.get(G.F_Z), 0.0)
    Aa, Ab, Aab = apply_a(a), apply_a(b), apply_a(a + 2.0 * b)
    scale = float(np.abs(Aab).max())
    assert float(np.abs(Aab - (Aa + 2.0 * Ab)).max()) <= 1e-12 * scale
    sab, sba = float((a * Ab).sum()), float((b * Aa).sum())
    assert abs(sab - sba) <= 1e-10 * max(abs(sab), 1.0)
    assert float((a * Aa).sum()) > 0.0                      # positive definite on the fluid cells
    Ma, Mb = apply_m(a), apply_m(b)
    mab, mba = float((a * Mb).sum()), float((b * Ma).sum())
    assert abs(mab - mba) <= 1e-10 * max(abs(mab), 1.0)
    assert float((a * Ma).sum()) > 0.0
    g.close()


def test_16384_properties():
    """BASELINE config[4] size on one GPU (the bench workload): two sub-steps of basic-fill at
    16384^2.  The count plane sums to the marker count, the fluid block is still the 40 % x 50 %
    rectangle it started as (nothing leaks through the walls), the capped solve reduces the
    residual, pressure is non-negative and finite, the device footprint is what DESIGN.md says."""
    from euler_b200 import gpu as G
    n = 16384
    scn = Scenario(synthetic("basic-fill", n, n), n, n, row_major_markers=True)
    g = G.EulerGpu.from_scenario(scn, precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST, pcg_check_every=25)
    n0 = int(g.stats().n_markers)
    res = []
    for _ in range(2):
        g.substep(g.calculate_timestep(0.1))
        res.append(float(g.stats().last_residual))
    st = g.stats()
    assert st.last_iterations == 100 and np.isfinite(res).all()
    cnt = g.read_marker_count()
    assert int(cnt.astype(np.int64).sum()) == int(st.n_markers) == n0
    assert not cnt[scn.solid != 0].any()
    win = np.zeros((n, n), np.float64)
    g.read_window(G.F_P, 0, 0, n, 64, win)                   # the bottom rows: deepest water
    wet = cnt[:64] != 0
    assert wet.any() and float(win[:64][wet].min()) >= 0.0 and np.isfinite(win[:64]).all()
    assert 35e9 < int(st.device_bytes) < 50e9
    g.close()


@pytest.mark.parametrize("name,nx,ny,substeps", [("block", 1100, 200, 160), ("waterfall", 700, 260, 120)])
def test_tile_list_of_the_grid_stages_follows_the_fluid(name, nx, ny, substeps):
    """The grid stages stream only the 512x32-cell tiles that held or bordered fluid within the last
    three sub-steps (common.cuh GridTiles).  With the iteration cap at 0 on both sides nothing
    depends on summation order, so a body of fluid falling through several tile rows must stay
    bit-identical to the oracle in EVERY plane — also where the fluid has left (tiles dropped from
    the list must have been cleared first) and where it arrives (tiles picked up in time)."""
    from euler_b200 import gpu as G
    from oracle.oracle import Oracle
    text = resample(shipped_text(name), nx - 2, ny - 2)
    o = Oracle(nx, ny, text)
    o.c.precon_mode = 1; o.c.quirk_marker_dt_leak = 0; o.c.max_iterations = 0
    g = G.EulerGpu.from_scenario(Scenario(text, nx, ny), precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST,
                                 max_iterations=0)
    cells_seen = set()
    y_first = None
    for i in range(substeps):
        dt = o.calculate_timestep(0.1)
        assert g.calculate_timestep(0.1) == dt
        o.substep(dt); g.substep(dt)
        cells_seen.add(int(g.stats().grid_cells))
        if i % 20 == 19 or i == substeps - 1:
            assert same_bits(g.get(G.F_COUNT), o.count), i
            assert same_bits(g.get(G.F_PREV_COUNT), o.prev_count), i
            assert same_bits(g.get(G.F_U), o.u) and same_bits(g.get(G.F_V), o.v), i
            assert same_bits(g.get(G.F_UTMP), o.utmp) and same_bits(g.get(G.F_VTMP), o.vtmp), i
            assert same_bits(g.get(G.F_MARKERS), o.markers), i
            assert int(g.stats().rng_state) == int(o.c.rng_state)
        rows = np.nonzero((o.count != 0).any(axis=1))[0]
        if y_first is None and len(rows):
            y_first = (int(rows.min()), int(rows.max()))
    rows = np.nonzero((o.count != 0).any(axis=1))[0]
    assert len(cells_seen) > 2, "the tile list never changed: the case does not exercise it"
    assert min(cells_seen) < nx * ny, "the grid stages streamed every tile all the time"
    assert (int(rows.min()) // 32, int(rows.max()) // 32) != (y_first[0] // 32, y_first[1] // 32), "fluid stayed in its tile rows"
    g.close()
