"""Mixed-precision PCG (euler_params.pcg_dtype = FP32): what the mode is allowed
to change, checked on the CPU mirror (oracle pcg_mixed) against the fp64 red-black solve from
IDENTICAL pre-projection state.

The mode is not in the reference, so there is no reference output to pin it to; the bar is
(1) converged solves give the fp64 solve's answer to fp32 rounding of the velocities
    (stated tolerance: 2e-7 relative in u, v and p — measured <= 8e-8 / 6e-8 on the shipped
    scenarios), in at most a few more iterations;
(2) without the residual replacement the same solves drift further (that is why it is there);
(3) a solve cut off at the iteration cap ends with a residual of the same size as the fp64 one.
The GPU kernels are checked against this mirror in tests/test_gpu_mixed.py."""
import numpy as np
import pytest

from euler_b200 import shipped_text, resample
from oracle.oracle import Oracle, PRECON_REDBLACK, PCG_FP32


def _pair(name, nx, ny, refresh):
    text = shipped_text(name) if (nx, ny) == (100, 40) else resample(shipped_text(name), nx - 2, ny - 2)
    o = Oracle(nx, ny, text); o.c.precon_mode = PRECON_REDBLACK; o.c.quirk_marker_dt_leak = 0
    m = Oracle(nx, ny, text); m.c.precon_mode = PRECON_REDBLACK; m.c.quirk_marker_dt_leak = 0
    m.c.pcg_dtype = PCG_FP32; m.c.refresh_every = refresh
    return o, m


def _compare_solves(name, nx, ny, frames, refresh):
    """fp64 trajectory; at every sub-step the mixed mirror repeats project() from the same
    utmp/vtmp/count.  Returns worst relative deviations and iteration counts."""
    o, m = _pair(name, nx, ny, refresh)
    worst = {"u": 0.0, "p": 0.0, "extra_it": 0, "solves": 0, "capped": 0, "res_ratio": 0.0}
    for _ in range(frames):
        left, steps = np.float32(0.1), 0
        while left > 0 and steps < 8:
            dt = o.calculate_timestep(float(left)); left = np.float32(left - np.float32(dt)); steps += 1
            o.substep(dt)
            for plane in ("utmp", "vtmp", "count", "prev_count", "adiag"):
                getattr(m, plane)[...] = getattr(o, plane)
            m.project(dt)
            assert m.c.last_solve_skipped == o.c.last_solve_skipped
            if o.c.last_solve_skipped:
                continue
            worst["solves"] += 1
            if o.c.last_iterations >= o.c.max_iterations:
                worst["capped"] += 1
                worst["res_ratio"] = max(worst["res_ratio"], m.c.last_residual / max(o.c.last_residual, 1e-30))
                continue
            su = max(1.0, float(np.abs(o.u).max()), float(np.abs(o.v).max()))
            sp = max(1.0, float(np.abs(o.p).max()))
            worst["u"] = max(worst["u"], float(np.abs(m.u - o.u).max()) / su, float(np.abs(m.v - o.v).max()) / su)
            worst["p"] = max(worst["p"], float(np.abs(m.p - o.p).max()) / sp)
            worst["extra_it"] = max(worst["extra_it"], m.c.last_iterations - o.c.last_iterations)
            assert m.c.last_residual <= m.c.tol
    return worst


@pytest.mark.parametrize("name", ["basic", "block", "filter", "waterfall", "weird-edges"])
def test_converged_solves_match_fp64_to_fp32_rounding(name):
    w = _compare_solves(name, 100, 40, 40, refresh=10)
    assert w["solves"] > 20
    assert w["u"] <= 2e-7 and w["p"] <= 2e-7, w
    assert w["extra_it"] <= 20, w


def test_residual_replacement_is_what_buys_the_accuracy():
    with_r = _compare_solves("waterfall", 100, 40, 60, refresh=10)
    without = _compare_solves("waterfall", 100, 40, 60, refresh=0)
    assert with_r["p"] <= 2e-7
    assert without["p"] > 10 * with_r["p"], (with_r, without)
    assert without["u"] <= 1e-5          # still inside north_star's velocity tolerance


def test_capped_solves_end_with_a_comparable_residual():
    """weird-edges at 192x192 hits the 100-iteration cap: both solves are
    unconverged iterates; the mixed one's residual is not worse than 4x the fp64 one's."""
    w = _compare_solves("weird-edges", 192, 192, 12, refresh=10)
    assert w["capped"] >= 5, w
    assert w["res_ratio"] <= 4.0, w


def test_fp32_preconditioner_is_the_rounded_fp64_one():
    o, m = _pair("block", 100, 40, 10)
    for _ in range(20):
        o.step_frame()
    dt = o.calculate_timestep(0.1)
    o.substep(dt)
    for plane in ("count", "adiag"):
        getattr(m, plane)[...] = getattr(o, plane)
    fl = o.count != 0
    o.r[:] = o.b
    o.apply_preconditioner(o.r, o.z)
    m.r32[:] = o.b.astype(np.float32)
    m.rb_build32(); m.rb_apply32(m.r32, m.z32)
    assert np.array_equal(m.pc32[fl], o.precon[fl].astype(np.float32))
    scale = float(np.abs(o.z[fl]).max())
    assert float(np.abs(m.z32[fl] - o.z[fl]).max()) <= 1e-6 * scale
    # A s in fp32 vs fp64 on the same (fp32-representable) s
    s = np.random.default_rng(5).standard_normal(o.s.shape).astype(np.float32)
    out32 = np.zeros_like(s); out64 = np.zeros(s.shape)
    m.apply_a32(s, out32); o.apply_a(s.astype(np.float64), out64)
    assert float(np.abs(out32[fl] - out64[fl]).max()) <= 4e-6
