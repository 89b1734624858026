"""The GPU-parallel red-black mode against the REFERENCE'S CONVERGED ANSWER (north_star: "a
GPU-parallel red-black/multicolour variant that converges to the same residual tolerance").

The reference caps PCG at 100 iterations (main.c:735) and at scale stops far from converged, where
any two preconditioners (or summation orders) give different iterates.  The comparison that means
something is the one SURVEY §7 (hard part 2) prescribes: raise the cap ON BOTH SIDES until
||r||inf <= 1e-6f (main.c:736, 756) and compare the solutions.  Oracle side = the reference's own
algorithm (natural-order IC(0), sequential dot products: oracle/euler_oracle.c, pinned bit for
bit to the unmodified reference); GPU side = red-black IC(0), tree dot products.

From IDENTICAL pre-projection state (the oracle's utmp, vtmp and count plane) both must reach the
tolerance, and then: p within 1e-5 of max|p|, u and v within 1e-5 of their maxima (north_star's
"1e-5 in float"), and the same post-projection divergence norm (the reference's, not zero: p is
clamped at 0, main.c:773-779)."""
import numpy as np
import pytest

from conftest import SCENARIOS
from euler_b200 import Scenario, shipped_text, synthetic

pytestmark = pytest.mark.gpu

CAP = 20000


def _to_projection(o):
    """The stages of one sub-step up to project() (main.c:852-889); returns dt."""
    dt = o.calculate_timestep(0.1)
    o.advect_markers(dt); o.refresh_marker_counts(); o.update_fluid_sources()
    o.extrapolate(o.u, 1); o.extrapolate(o.v, 2); o.zero_bounds(o.u, 1); o.zero_bounds(o.v, 2)
    o.advect_u(dt); o.advect_v(dt); o.apply_body_forces(dt)
    o.zero_bounds(o.utmp, 1); o.zero_bounds(o.vtmp, 2)
    return dt


def _div_norm(u, v, fl):
    d = (u[1:-1, 1:-1] - u[1:-1, :-2]) + (v[1:-1, 1:-1] - v[:-2, 1:-1])      # main.c:720, h = 1
    m = fl[1:-1, 1:-1]
    return float(np.abs(d[m]).max()) if m.any() else 0.0


def _compare_projection(o, g, G, dt, label):
    """project() on both sides from the oracle's pre-projection state; returns the iteration counts."""
    g.set(G.F_UTMP, o.utmp); g.set(G.F_VTMP, o.vtmp); g.set(G.F_COUNT, o.count)
    o.project(dt)
    g.run_stage(G.S_PROJECT, dt)
    st = g.stats()
    fl = o.count != 0
    if o.c.last_solve_skipped:
        assert st.last_iterations == 0, label
    else:
        assert 0 < o.c.last_iterations < CAP and o.c.last_residual <= 1e-6, (label, "reference algorithm did not converge")
        assert 0 < st.last_iterations < CAP and st.last_residual <= 1e-6, (label, "red-black did not converge")
    p = g.get(G.F_P)
    if fl.any():
        assert float(np.abs(p[fl] - o.p[fl]).max()) <= 1e-5 * max(1.0, float(np.abs(o.p[fl]).max())), label
    gu, gv = g.get(G.F_U), g.get(G.F_V)
    for got, ref, what in ((gu, o.u, "u"), (gv, o.v, "v")):
        assert float(np.abs(got - ref).max()) <= 1e-5 * max(1.0, float(np.abs(ref).max())), (label, what)
    scale = max(1.0, float(np.abs(o.u).max()), float(np.abs(o.v).max()))
    assert abs(_div_norm(gu, gv, fl) - _div_norm(o.u, o.v, fl)) <= 2e-5 * scale, (label, "divergence norm")
    return int(o.c.last_iterations), int(st.last_iterations)


@pytest.mark.parametrize("name", SCENARIOS)
def test_red_black_converged_equals_reference_converged(name):
    """All five shipped scenarios, 100x40, projections of sub-steps spread over 40 frames (free
    fall, first impact, sloshing, sources running)."""
    from euler_b200 import gpu as G
    from oracle.oracle import Oracle
    text = shipped_text(name)
    o = Oracle(100, 40, text)
    o.c.precon_mode = 0                      # the reference's algorithm
    o.c.max_iterations = CAP
    g = G.EulerGpu.from_scenario(Scenario(text, 100, 40), precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST,
                                 max_iterations=CAP, pcg_check_every=50)
    solved = 0
    totals = [0, 0]
    for sub in range(120):
        dt = _to_projection(o)
        if sub % 6 == 0 or sub < 4:
            a, b = _compare_projection(o, g, G, dt, "%s sub-step %d" % (name, sub))
            solved += 1 if a else 0
            totals[0] += a; totals[1] += b
        else:
            o.project(dt)
    assert solved >= 5, "too few active solves compared"
    print("%s: %d solves compared, iterations to 1e-6: reference IC(0) %d, red-black %d" % (name, solved, *totals))
    g.close()


def test_red_black_converged_equals_reference_converged_1024():
    """1024^2 basic-fill (hydrostatic column: a smooth right-hand side, the slowest kind to
    converge): the reference's algorithm needs several hundred iterations here and stops at 100;
    with the cap raised both reach 1e-6 and agree."""
    from euler_b200 import gpu as G
    from oracle.oracle import Oracle
    n = 1024
    text = synthetic("basic-fill", n, n)
    o = Oracle(n, n, text)
    o.c.precon_mode = 0
    o.c.max_iterations = CAP
    g = G.EulerGpu.from_scenario(Scenario(text, n, n), precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST,
                                 max_iterations=CAP, pcg_check_every=50)
    dt = _to_projection(o)
    a, b = _compare_projection(o, g, G, dt, "basic-fill 1024 sub-step 0")
    assert a > 100, "this case is meant to need more than the reference's cap"
    print("basic-fill 1024^2: iterations to 1e-6: reference IC(0) %d, red-black %d" % (a, b))
    g.close()
