"""The oracle (oracle/euler_oracle.c, our CPU restatement) against the reference.

Pins: (1) known-answer vectors generated from the UNMODIFIED reference
(tests/golden/known_answers.json, generator committed next to it; same quantities as
BASELINE.md §3); (2) when oracle/_ref is built, bit-for-bit equality of every plane and of the
marker ARRAY with the reference after whole frames and after each single stage."""
import numpy as np
import pytest

from conftest import SCENARIOS, same_bits, load_state
from euler_b200 import shipped_text, resample
from oracle.oracle import Oracle, Reference, ref_available, fnv1a, U, V

FRAMES = [0, 1, 10, 50]


def snapshot(o):
    return {"markers": o.n_markers, "fluid_cells": int((o.count != 0).sum()),
            "fnv_count": "%016x" % fnv1a(o.count),
            "sum_abs_u": float(np.abs(o.u.astype(np.float64)).sum()),
            "sum_abs_v": float(np.abs(o.v.astype(np.float64)).sum()),
            "rng_state": "%016x" % int(o.c.rng_state)}


@pytest.mark.parametrize("name", SCENARIOS)
def test_known_answers_shipped_scenarios(name, known_answers):
    o = Oracle(100, 40, shipped_text(name))
    frame = 0
    for target in FRAMES:
        while frame < target:
            o.step_frame()
            frame += 1
        assert snapshot(o) == known_answers["%s@100x40/f%d" % (name, target)], (name, target)


def test_known_answers_resampled(known_answers):
    o = Oracle(64, 48, resample(shipped_text("block"), 62, 46))
    frame = 0
    for target in (0, 5, 20):
        while frame < target:
            o.step_frame()
            frame += 1
        assert snapshot(o) == known_answers["block@64x48/f%d" % target]


def test_baseline_md_hashes():
    """Two of the hashes listed in BASELINE.md §3 (survey probe of the reference)."""
    o = Oracle(100, 40, shipped_text("block"))
    assert "%016x" % fnv1a(o.count) == "02c0252f3a589ac3" and o.n_markers == 4488
    for _ in range(10):
        o.step_frame()
    assert "%016x" % fnv1a(o.count) == "02ca8c16b9f88d9f"


def _assert_same_state(o, r, what):
    assert o.n_markers == r.n_markers, what
    assert same_bits(o.markers, r.markers), what + ": marker array"
    for f in ("u", "v", "utmp", "vtmp", "count", "prev_count", "precon"):
        assert same_bits(getattr(o, f), getattr(r, f)), "%s: %s" % (what, f)
    assert int(o.c.rng_state) == r.rng_state, what + ": rng"


@pytest.mark.skipif(not ref_available(100, 40), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("name", SCENARIOS)
def test_bit_identical_to_reference_frames(name):
    text = shipped_text(name)
    o, r = Oracle(100, 40, text), Reference(100, 40)
    r.init_from_text(text)
    _assert_same_state(o, r, "init")
    for f in range(30):
        o.step_frame()
        r.step_frame()
        _assert_same_state(o, r, "%s frame %d" % (name, f + 1))


@pytest.mark.skipif(not ref_available(256, 256), reason="oracle/_ref not built")
def test_bit_identical_to_reference_256():
    text = resample(shipped_text("weird-edges"), 254, 254)
    o, r = Oracle(256, 256, text), Reference(256, 256)
    r.init_from_text(text)
    for f in range(4):
        o.step_frame()
        r.step_frame()
        _assert_same_state(o, r, "frame %d" % (f + 1))
    assert o.c.last_iterations == 100          # the reference's cap is hit at this size


def _from_state(cls_obj, st):
    cls_obj.u[:] = st["u"]; cls_obj.v[:] = st["v"]
    cls_obj.count[:] = st["count"]; cls_obj.prev_count[:] = st["prev_count"]
    cls_obj.precon[:] = st["precon"]
    cls_obj.set_markers(st["markers"])


@pytest.mark.skipif(not ref_available(100, 40), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["block", "waterfall", "weird-edges"])
def test_each_stage_against_reference(name):
    """Every stage run from IDENTICAL input state (the golden state after 10 frames)."""
    text = shipped_text(name)
    st = load_state(name)
    o, r = Oracle(100, 40, text), Reference(100, 40)
    r.init_from_text(text)
    _from_state(o, st); _from_state(r, st)
    o.c.rng_state = int(st["rng_state"]); r.rng_state = int(st["rng_state"])
    dt_o, dt_r = o.calculate_timestep(0.1), r.calculate_timestep(0.1)
    assert dt_o == dt_r
    dt = dt_o
    o.advect_markers(dt); r.advect_markers(dt)
    assert same_bits(o.markers, r.markers)
    o.refresh_marker_counts(); r.refresh_marker_counts()
    assert same_bits(o.count, r.count) and same_bits(o.prev_count, r.prev_count)
    assert same_bits(o.markers, r.markers)
    o.update_fluid_sources(); r.update_fluid_sources()
    assert same_bits(o.markers, r.markers) and same_bits(o.count, r.count)
    assert int(o.c.rng_state) == r.rng_state
    for t, f in ((U, "u"), (V, "v")):
        o.extrapolate(getattr(o, f), t); r.extrapolate(getattr(r, f), t)
    for t, f in ((U, "u"), (V, "v")):
        o.zero_bounds(getattr(o, f), t); r.zero_bounds(getattr(r, f), t)
    assert same_bits(o.u, r.u) and same_bits(o.v, r.v)
    o.advect_u(dt); r.advect_u(dt); o.advect_v(dt); r.advect_v(dt)
    o.apply_body_forces(dt); r.apply_body_forces(dt)
    o.zero_bounds(o.utmp, U); r.zero_bounds(r.utmp, U)
    o.zero_bounds(o.vtmp, V); r.zero_bounds(r.vtmp, V)
    assert same_bits(o.utmp, r.utmp) and same_bits(o.vtmp, r.vtmp)
    # preconditioner and A applied to the same vectors
    rng = np.random.default_rng(5)
    vec = rng.standard_normal((40, 100))
    o.build_rhs(dt)
    r.adiag[:] = o.adiag
    zo, zr = np.zeros((40, 100)), np.zeros((40, 100))
    o.apply_preconditioner(vec, zo); r.apply_preconditioner(vec, zr)
    fl = o.count != 0
    assert same_bits(zo[fl], zr[fl]) and same_bits(o.precon, r.precon) and same_bits(o.q, r.q)
    ao, ar = np.zeros((40, 100)), np.zeros((40, 100))
    o.apply_a(vec, ao); r.apply_a(vec, ar)
    assert same_bits(ao[fl], ar[fl])
    assert o.dot(vec, zo) == r.dot(vec, zr)
    o.project(dt); r.project(dt)
    assert same_bits(o.u, r.u) and same_bits(o.v, r.v)


def test_marker_dt_carry_over_quirk_is_observable():
    """main.c:464,501,518: `dt -= t_prev` leaks into the following markers.  The oracle
    reproduces it by default; switching it off changes the trajectory."""
    text = shipped_text("filter")         # first fires in frame 29 here (block: frame 275)
    a, b = Oracle(100, 40, text), Oracle(100, 40, text)
    b.c.quirk_marker_dt_leak = 0
    differs = False
    for _ in range(40):
        a.step_frame(); b.step_frame()
        if a.n_markers != b.n_markers or not same_bits(a.markers, b.markers):
            differs = True
            break
    assert differs


def test_source_exhaustion_latch():
    """MAX_MARKER_COUNT-1 latch of update_fluid_sources (main.c:281, 290)."""
    o = Oracle(16, 12, "????????\n????????\n")
    cap = 4 * 16 * 12 - 1
    o.set_markers(np.tile(np.array([[8.5, 5.5]], np.float32), (cap - 3, 1)))
    o.count[:] = 0
    o.update_fluid_sources()
    assert o.n_markers == cap and o.c.source_exhausted == 1
    o.count[:] = 0
    o.update_fluid_sources()
    assert o.n_markers == cap


# ---- --rainbow colour transport ------------------------------------------------

def _rainbow_snapshot(o):
    fl = o.count != 0
    out = {"fluid_cells": int(fl.sum()), "fnv_count": "%016x" % fnv1a(o.count)}
    for k, plane in (("r", o.cr), ("g", o.cg), ("b", o.cb)):
        masked = np.where(fl, plane, np.float32(0)).astype(np.float32)
        out["fnv_" + k] = "%016x" % fnv1a(masked.view(np.uint8))
    return out


@pytest.mark.parametrize("name", SCENARIOS)
def test_rainbow_known_answers(name):
    """The restated colour transport against known answers produced by the UNMODIFIED reference
    with g_rainbow_enabled (tests/golden/make_golden.py)."""
    import json, os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "rainbow_answers.json")) as f:
        want = json.load(f)
    o = Oracle(100, 40, rainbow=True)
    o.init_from_text(shipped_text(name))
    frame = 0
    for target in (0, 1, 10, 30):
        while frame < target:
            o.step_frame()
            frame += 1
        assert _rainbow_snapshot(o) == want["%s@100x40/f%d" % (name, target)], (name, target)


@pytest.mark.skipif(not ref_available(100, 40), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("name", ["block", "waterfall", "weird-edges"])
def test_rainbow_planes_bit_identical_to_reference(name):
    """Whole g_r/g_g/g_b planes (stale values at non-fluid cells included: the reference copies
    the whole tmp plane back, main.c:875) after every frame, and the planes of an `r` key press
    (colorize on the evolved fluid, main.c:971-974)."""
    text = shipped_text(name)
    o = Oracle(100, 40, rainbow=True); o.init_from_text(text)
    r = Reference(100, 40); r.init_from_text(text, rainbow=True)
    for f in range(25):
        assert o.step_frame() is not None
        r.step_frame()
        for a, b, what in ((o.cr, r.cr, "r"), (o.cg, r.cg, "g"), (o.cb, r.cb, "b"), (o.count, r.count, "count")):
            assert same_bits(a, b), (name, f, what)
    o.colorize(); r.L.colorize()
    for a, b in ((o.cr, r.cr), (o.cg, r.cg), (o.cb, r.cb)):
        assert same_bits(a, b)


@pytest.mark.skipif(not (ref_available(100, 40) and ref_available(64, 48)), reason="oracle/_ref not built")
def test_sim_init_fuzz_against_reference():
    """The parser / ring of sinks / marker seeding of sim_init (main.c:209-274) on random texts —
    junk characters, ragged, over-long and empty lines, missing final newline — and the first
    frames that follow: the restatement equals the unmodified reference bit for bit."""
    rng = np.random.default_rng(20261017)
    alphabet = np.array(list("X0?=   ab"))
    for case in range(24):
        nx, ny = ((100, 40), (64, 48))[case % 2]
        lines = ["".join(rng.choice(alphabet, size=int(rng.integers(0, nx + 30)))) for _ in range(int(rng.integers(0, ny + 8)))]
        text = "\n".join(lines) + ("\n" if case % 3 else "")
        o, r = Oracle(nx, ny, text), Reference(nx, ny)
        r.init_from_text(text)
        _assert_same_state(o, r, "fuzz %d init" % case)
        for f in range(8):
            o.step_frame()
            r.step_frame()
            _assert_same_state(o, r, "fuzz %d frame %d" % (case, f + 1))
