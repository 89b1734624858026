"""The in-kernel timeline of the PCG iteration kernels (EULER_TRACE, euler_gpu_trace_read,
include/euler_gpu.h): diagnostics of the new build (the reference has
misc/debug.c).  Off by default; when on, every launch of the two iteration kernels leaves one
slot whose timestamps are ordered, and the results of the solve do not change."""
import numpy as np
import pytest

from conftest import same_bits
from euler_b200 import Scenario, shipped_text

pytestmark = pytest.mark.gpu


def _run(monkeypatch, slots):
    from euler_b200 import gpu as G
    if slots:
        monkeypatch.setenv("EULER_TRACE", str(slots))
    else:
        monkeypatch.delenv("EULER_TRACE", raising=False)
    scn = Scenario(shipped_text("weird-edges"), 100, 40)      # iterates from the first frame on
    g = G.EulerGpu.from_scenario(scn, precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST)
    for _ in range(4):
        g.step_frame()
    tr = g.trace_read(2048)
    again = g.trace_read(2048)
    out = (tr, again, g.get(G.F_U), g.get(G.F_V), g.read_marker_count(), g.stats().pcg_iterations)
    g.close()
    return out


def test_trace_is_off_by_default(monkeypatch):
    tr, again, *_ = _run(monkeypatch, 0)
    assert len(tr) == 0 and len(again) == 0


def test_trace_slots_are_ordered_and_do_not_change_the_solve(monkeypatch):
    tr, again, u, v, count, iters = _run(monkeypatch, 1024)
    _, _, u0, v0, count0, iters0 = _run(monkeypatch, 0)
    assert iters == iters0 > 0
    assert same_bits(u, u0) and same_bits(v, v0) and same_bits(count, count0)
    assert len(again) == 0, "reading starts the recording over"
    assert 0 < len(tr) <= 1024
    # launches enqueued past convergence return at once and leave their slot empty
    tr = tr[tr[:, 8] != 0]
    kinds = tr[:, 8].astype(np.int64)
    assert set(kinds.tolist()) == {1, 2}, "search+apply and tail launches"
    assert np.all(kinds[:-1] != kinds[1:]), "a solve alternates the two kernels and ends with a tail"
    lo = lambda k: (~tr[:, 2 * k]).astype(np.int64)      # minima are stored complemented
    hi = lambda k: tr[:, 2 * k + 1].astype(np.int64)
    for k in range(4):
        assert np.all(lo(k) <= hi(k)), "pair %d: first block not after last block" % k
    assert np.all(lo(0) <= hi(3)), "a kernel exits after it starts"
    assert np.all(lo(2) <= hi(3)), "blocks leave after their rows are done"
    # launches are serialised on the stream: a kernel starts after the previous one's last exit
    assert np.all(lo(0)[1:] >= hi(3)[:-1])
