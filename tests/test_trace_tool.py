"""tools/iter_trace.py turns the slots of euler_gpu_trace_read (include/euler_gpu.h) into the phase
table of DESIGN.md section 4: checked here on a synthetic timeline (no GPU)."""
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _tool():
    spec = importlib.util.spec_from_file_location("iter_trace", os.path.join(_HERE, "..", "tools", "iter_trace.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _slot(kind, t0, ramp, collect, rows_first, rows_last, epilogue):
    """One launch: first block starts at t0 ns, the last `ramp` later; scalars known `collect` after
    each start; rows take rows_first / rows_last; the kernel exits `epilogue` after the last rows."""
    inv = lambda v: np.uint64(~np.uint64(v))
    s = np.zeros(16, dtype=np.uint64)
    st0, st1 = t0, t0 + ramp
    c0, c1 = st0 + collect, st1 + collect
    w0, w1 = c0 + rows_first, c1 + rows_last
    e0, e1 = w0 + 100, w1 + epilogue
    s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7] = inv(st0), st1, inv(c0), c1, inv(w0), w1, inv(e0), e1
    s[8] = kind
    return s, e1


def test_phase_table_of_a_synthetic_timeline():
    m = _tool()
    slots, t = [], 1_000_000
    for i in range(60):
        for kind, rows in ((1, 40_000), (2, 80_000)):
            s, end = _slot(kind, t + 4_000, 300, 1_500, rows - 10_000, rows, 4_200)
            slots.append(s)
            t = end
        if i == 30:                                    # a launch past convergence leaves an empty slot
            slots.append(np.zeros(16, dtype=np.uint64))
    res = m.analyse(np.array(slots))
    for kind, rows in ((1, 40.0), (2, 80.0)):
        r = res[kind]
        assert abs(r["gap"] - 4.0) < 1e-9 and abs(r["ramp"] - 0.3) < 1e-9
        assert abs(r["collect_first"] - 1.5) < 1e-9 and abs(r["collect_last"] - 1.5) < 1e-9
        assert abs(r["rows_last"] - rows) < 1e-9 and abs(r["rows_first"] - (rows - 10.0)) < 1e-9
        assert abs(r["end_spread"] - 10.3) < 1e-9 and abs(r["epilogue"] - 4.2) < 1e-9
        assert abs(r["period"] - (4.0 + 0.3 + 1.5 + rows + 4.2)) < 1e-9
