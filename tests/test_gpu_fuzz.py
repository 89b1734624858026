"""GPU parity on RANDOM scenarios (the generator of tests/test_oracle.py's fuzz, where the
oracle is pinned to the unmodified reference on the same texts): sources, sinks and solids in
arbitrary places, ragged and over-long lines.  Reference-faithful mode with reference-order dot
products must stay bit-identical to the oracle over whole frames; the red-black mode keeps the
classification bit-exact and the velocities within 1e-5."""
import numpy as np
import pytest

from conftest import same_bits
from euler_b200 import Scenario

pytestmark = pytest.mark.gpu


def fuzz_text(case):
    """Case `case` of the shared generator (same stream as test_sim_init_fuzz_against_reference)."""
    rng = np.random.default_rng(20261017)
    alphabet = np.array(list("X0?=   ab"))
    for c in range(case + 1):
        nx, ny = ((100, 40), (64, 48))[c % 2]
        lines = ["".join(rng.choice(alphabet, size=int(rng.integers(0, nx + 30)))) for _ in range(int(rng.integers(0, ny + 8)))]
        text = "\n".join(lines) + ("\n" if c % 3 else "")
    return text, nx, ny


@pytest.mark.parametrize("case", [0, 3, 6, 7, 8, 14, 16, 17])
def test_random_scenarios_bit_identical_in_reference_order_mode(case):
    from euler_b200 import gpu as G
    from oracle.oracle import Oracle
    text, nx, ny = fuzz_text(case)
    o = Oracle(nx, ny, text)
    o.c.quirk_marker_dt_leak = 0
    g = G.EulerGpu.from_scenario(Scenario(text, nx, ny), precon=G.PRECON_IC0_WAVEFRONT, dot_mode=G.DOT_REFERENCE_ORDER,
                                 marker_mode=G.MARKERS_FAST)
    for f in range(6):
        assert o.step_frame() == g.step_frame()
        assert same_bits(g.get(G.F_COUNT), o.count), "frame %d" % f
    assert same_bits(g.get(G.F_MARKERS), o.markers)
    assert same_bits(g.get(G.F_U), o.u) and same_bits(g.get(G.F_V), o.v)
    st = g.stats()
    assert st.pcg_iterations == o.c.total_iterations and st.solves == o.c.total_solves
    assert int(st.rng_state) == int(o.c.rng_state)
    g.close()


@pytest.mark.parametrize("case", [4, 5, 12])
def test_random_scenarios_red_black(case):
    from euler_b200 import gpu as G
    from oracle.oracle import Oracle, PRECON_REDBLACK
    text, nx, ny = fuzz_text(case)
    o = Oracle(nx, ny, text)
    o.c.quirk_marker_dt_leak = 0
    o.c.precon_mode = PRECON_REDBLACK
    g = G.EulerGpu.from_scenario(Scenario(text, nx, ny), precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST)
    for f in range(6):
        assert o.step_frame() == g.step_frame()
        assert same_bits(g.get(G.F_COUNT), o.count), "frame %d" % f
    for fld, ref in ((G.F_U, o.u), (G.F_V, o.v)):
        assert float(np.abs(g.get(fld) - ref).max()) <= 1e-5 * max(1.0, float(np.abs(ref).max()))
    g.close()


def test_host_program_exports_its_final_state_as_a_scenario(tmp_path):
    """bin/euler-gpu --export: the file is the final state in the scenario format —
    parsing it gives the static masks back and, as fluid, the cells the count plane marks wet
    (hash of the count plane printed by the same run == hash of the oracle's)."""
    import json
    import os
    import subprocess
    from conftest import ROOT
    from euler_b200 import shipped_text
    from oracle.oracle import Oracle, fnv1a
    exe = os.path.join(ROOT, "bin", "euler-gpu")
    assert os.path.exists(exe), "run `make host`"
    src = tmp_path / "waterfall.txt"
    src.write_bytes(shipped_text("waterfall"))
    dst = tmp_path / "exported.txt"
    out = subprocess.run([exe, "--headless", "--frames", "20", "--exact-dot", "--export", str(dst), str(src)],
                         check=True, capture_output=True, text=True, cwd=ROOT).stdout
    res = json.loads(out.strip().splitlines()[-1])
    o = Oracle(100, 40, shipped_text("waterfall"))          # reference-faithful defaults, like the CLI
    for _ in range(20):
        o.step_frame()
    assert res["fnv_count"] == "%016x" % fnv1a(o.count)
    s = Scenario(dst.read_bytes(), 100, 40)
    assert np.array_equal(s.solid, o.solid) and np.array_equal(s.source, o.source) and np.array_equal(s.sink, o.sink)
    expect = ((o.count != 0) & (o.solid == 0) & (o.sink == 0)) | (o.source != 0)
    assert np.array_equal(s.fluid != 0, expect)
