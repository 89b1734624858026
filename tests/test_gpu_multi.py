"""Row-slab decomposition across 2, 4 and 8 GPUs against the single-GPU run of the
same scenario and against the oracle: NCCL halo exchange, cross-slab marker migration
(reference semantics at stake: main.c:464-537), distributed PCG scalars (main.c:629-667), the
cross-rank order of the source RNG draws (main.c:284-291).  Each case needs as many visible GPUs
as it has slabs (skipped otherwise); the logs of the 2-, 4- and 8-GPU runs on hardware are kept
under profiles/."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, same_bits
from euler_b200 import Scenario, shipped_text, resample

pytestmark = pytest.mark.gpu


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _exchange_blobs(out_dir, rank, nranks, blob):
    """all-gather of the IPC blobs through files (the side channel is the caller's business)"""
    import time
    with open(os.path.join(out_dir, "blob%d.tmp" % rank), "wb") as f:
        f.write(blob)
    os.rename(os.path.join(out_dir, "blob%d.tmp" % rank), os.path.join(out_dir, "blob%d.bin" % rank))
    blobs = []
    for r in range(nranks):
        path = os.path.join(out_dir, "blob%d.bin" % r)
        t0 = time.time()
        while not os.path.exists(path):
            if time.time() - t0 > 60:
                raise RuntimeError("peer blob missing")
            time.sleep(0.01)
        with open(path, "rb") as f:
            blobs.append(f.read())
    return blobs


def _worker(rank, nranks, uid, text, nx, ny, frames, out_dir, p2p=False, reinit=False, weighted=False):
    sys.path.insert(0, ROOT)
    from euler_b200 import gpu as G
    scn = Scenario(text, nx, ny)
    if weighted:
        weight = scn.fluid.sum(axis=1, dtype=np.uint64) * 100 + np.uint64(nx)
        row0, rows = G.slab_partition_weighted(weight, nranks, rank)
    else:
        row0, rows = G.slab_partition(ny, nranks, rank)
    g = G.EulerGpu.from_scenario(scn, precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST,
                                 device=rank, slab_row0=row0, slab_rows=rows)
    g.comm_init(rank, nranks, uid)
    if p2p:
        g.comm_p2p_import(_exchange_blobs(out_dir, rank, nranks, g.comm_p2p_export()))
    if reinit:
        # dirty the state, then sim_init again into the same handle (euler_gpu_reinit): the
        # run that follows must be the run of a fresh handle
        for _ in range(3):
            g.step_frame()
        g.reinit(scn.solid, scn.source, scn.sink, scn.markers, scn.rng_state)
    it0 = g.stats().pcg_iterations
    subs = [g.step_frame() for _ in range(frames)]
    st = g.stats()
    chk = g.check()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), row0=row0, rows=rows,
             count=g.get(G.F_COUNT), u=g.get(G.F_U), v=g.get(G.F_V), p=g.get(G.F_P),
             markers=g.get(G.F_MARKERS), subs=np.array(subs), iters=st.pcg_iterations - it0,
             rng=np.uint64(st.rng_state),
             chk_int=np.array([chk.n_markers, chk.fluid_cells, chk.count_sum, chk.count_hash], dtype=np.uint64),
             chk_sum=np.array([chk.sum_abs_u, chk.sum_abs_v, chk.sum_p]),
             chk_max=np.array([chk.max_abs_div, chk.max_abs_u, chk.max_abs_v]))
    g.close()


# (slabs, scenario, nx, ny, frames): 8 slabs of 100x64 / 4 of 100x40 are the degenerate geometry —
# 8 and 10 rows per slab, the minimum (2 x 4 halo rows) the decomposition accepts
_GRIDS = {2: [("block", 100, 40, 12), ("waterfall", 160, 96, 30), ("weird-edges", 256, 256, 5)],
          4: [("block", 100, 40, 12), ("waterfall", 160, 96, 30), ("weird-edges", 256, 256, 5)],
          8: [("block", 100, 64, 12), ("waterfall", 160, 96, 30), ("weird-edges", 256, 256, 5)]}
# (p2p, reinit, weighted): NCCL-only path, NVLink peer path, reinit into a used handle, weighted split.
# 2 slabs run the full matrix; 4 and 8 slabs (charged 4x / 8x on the GPU pool) the NVLink path on
# every grid plus one case each of the other modes.
_MODES_ALL = [(False, False, False), (True, False, False), (True, True, False), (True, False, True)]
_SLAB_CASES = [(2,) + g + m for g in _GRIDS[2] for m in _MODES_ALL]
for _n in (4, 8):
    _SLAB_CASES += [(_n,) + g + (True, False, False) for g in _GRIDS[_n]]
    _SLAB_CASES += [(_n,) + _GRIDS[_n][1] + (True, False, True), (_n,) + _GRIDS[_n][0] + (False, False, False),
                    (_n,) + _GRIDS[_n][2] + (True, True, False)]


@pytest.mark.parametrize("nranks,name,nx,ny,frames,p2p,reinit,weighted", _SLAB_CASES)
def test_slabs_match_single_gpu(nranks, name, nx, ny, frames, p2p, reinit, weighted, tmp_path):
    if _gpu_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    import torch.multiprocessing as mp
    from euler_b200 import gpu as G
    text = shipped_text(name)
    if (nx, ny) != (100, 40):
        text = resample(text, nx - 2, ny - 2)
    uid = G.comm_unique_id()
    mp.spawn(_worker, args=(nranks, uid, text, nx, ny, frames, str(tmp_path), p2p, reinit, weighted),
             nprocs=nranks, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(nranks)]

    ref = G.EulerGpu.from_scenario(Scenario(text, nx, ny), precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST)
    subs = [ref.step_frame() for _ in range(frames)]

    def merged(field, dtype):
        out = np.zeros((ny, nx), dtype)
        for p in parts:
            r0, n = int(p["row0"]), int(p["rows"])
            out[r0:r0 + n] = p[field][r0:r0 + n]
        return out
    assert all(list(p["subs"]) == subs for p in parts)
    assert sum(int(p["rows"]) for p in parts) == ny
    # cell classification: bit-exact
    assert same_bits(merged("count", np.uint8), ref.get(G.F_COUNT))
    # markers: same multiset (order across slabs is unspecified)
    m = np.concatenate([p["markers"] for p in parts])
    rm = ref.get(G.F_MARKERS)
    assert len(m) == len(rm)
    key = lambda a: a[np.lexsort((a[:, 0], a[:, 1]))]
    assert np.abs(key(m) - key(rm)).max() <= 1e-4
    # the RNG stream advanced identically on every rank and as on one GPU
    assert all(int(p["rng"]) == int(ref.stats().rng_state) for p in parts)
    for f, fld in (("u", G.F_U), ("v", G.F_V)):
        a, b = merged(f, np.float32), ref.get(fld)
        assert float(np.abs(a - b).max()) <= 1e-5 * max(1.0, float(np.abs(b).max())), f
    # euler_gpu_check (what bench.py prints at every N): integers add up exactly modulo 2^64,
    # sums / maxima agree to the tolerance of the fields themselves
    want = ref.check()
    got_int = np.zeros(4, dtype=np.uint64)
    for p in parts:
        got_int += p["chk_int"]
    assert [int(x) for x in got_int] == [want.n_markers, want.fluid_cells, want.count_sum, want.count_hash]
    got_sum = sum(p["chk_sum"] for p in parts)
    got_max = np.max([p["chk_max"] for p in parts], axis=0)
    for a, b in zip(got_sum, (want.sum_abs_u, want.sum_abs_v, want.sum_p)):
        assert abs(a - b) <= 1e-5 * max(1.0, abs(b))
    for a, b in zip(got_max[1:], (want.max_abs_u, want.max_abs_v)):
        assert abs(a - b) <= 1e-5 * max(1.0, abs(b))
    ref.close()


def _big_text(kind, n):
    from euler_b200 import synthetic
    return synthetic(kind, n, n) if kind == "basic-fill" else resample(shipped_text(kind), n - 2, n - 2)


def _worker_big(rank, nranks, uid, kind, n, steps, cap, out_dir):
    sys.path.insert(0, ROOT)
    from euler_b200 import gpu as G
    scn = Scenario(_big_text(kind, n), n, n, row_major_markers=True)
    weight = scn.fluid.sum(axis=1, dtype=np.uint64) * 4096 + np.uint64(max(1, n // 256))
    row0, rows = G.slab_partition_weighted(weight, nranks, rank)
    g = G.EulerGpu.from_scenario(scn, precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST,
                                 device=rank, slab_row0=row0, slab_rows=rows, pcg_check_every=50,
                                 max_iterations=cap)
    g.comm_init(rank, nranks, uid)
    g.comm_p2p_import(_exchange_blobs(out_dir, rank, nranks, g.comm_p2p_export()))
    for _ in range(steps):
        g.substep(g.calculate_timestep(0.1))
    st = g.stats()
    np.savez(os.path.join(out_dir, "big%d.npz" % rank), row0=row0, rows=rows,
             count=g.read_marker_count()[row0:row0 + rows], u=g.get(G.F_U)[row0:row0 + rows],
             v=g.get(G.F_V)[row0:row0 + rows], markers=np.uint64(st.n_markers), iters=st.pcg_iterations,
             rng=np.uint64(st.rng_state))
    g.close()


@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kind,n,steps,cap", [("weird-edges", 8192, 3, 100), ("basic-fill", 4096, 2, 100),
                                              ("basic-fill", 4096, 2, 20000)])
def test_large_grids_two_slabs(kind, n, steps, cap):
    """BASELINE config[3] geometry (weird-edges irregular solid mask at 8192^2 on 2 slabs: halo exchange and
    marker migration at scale) and a resting block at 4096^2 against the single-GPU run, weighted split,
    NVLink exchanges.  Marker totals, iteration counts and RNG state are equal.
      cap = 20000: both runs CONVERGE (||r||inf <= 1e-6, several thousand iterations) — the comparison
          that means something: u, v within 1e-5 of their maximum (north_star), count planes identical
          up to markers within one fp32 ulp of a cell edge.
      cap = 100 (the reference's, main.c:735): the solve stops far from converged (||r||inf ~ 10^3) and CG
          amplifies the different summation orders of the two runs' dot products; the iterates are
          different, equally unconverged ones (measured 2e-5 .. 2e-4 of max|v| depending on the reduction
          tree), so this case only checks that the runs stay on the same trajectory: 2e-3."""
    import tempfile
    import torch.multiprocessing as mp
    from euler_b200 import gpu as G
    uid = G.comm_unique_id()
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker_big, args=(2, uid, kind, n, steps, cap, tmp), nprocs=2, join=True)
        parts = [dict(np.load(os.path.join(tmp, "big%d.npz" % r))) for r in range(2)]
    ref = G.EulerGpu.from_scenario(Scenario(_big_text(kind, n), n, n, row_major_markers=True),
                                   precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST, pcg_check_every=50,
                                   max_iterations=cap)
    for _ in range(steps):
        ref.substep(ref.calculate_timestep(0.1))
    st = ref.stats()
    converged = cap > 100
    if converged:
        assert st.last_residual <= 1e-6 and st.last_iterations < cap
    full, fu, fv = ref.read_marker_count(), ref.get(G.F_U), ref.get(G.F_V)
    mismatched = 0
    for p in parts:
        r0, k = int(p["row0"]), int(p["rows"])
        mismatched += int((p["count"] != full[r0:r0 + k]).sum())
        assert int(p["rng"]) == int(st.rng_state)
        if converged:
            # thousands of iterations per solve: the two summation orders reach the tolerance within a
            # few iterations of each other (measured: 4 per solve of ~4150)
            assert abs(int(p["iters"]) - int(st.pcg_iterations)) <= max(2 * steps, int(0.005 * st.pcg_iterations))
        else:
            assert int(p["iters"]) == int(st.pcg_iterations)
        tol = 1e-5 if converged else 2e-3
        for a, b, what in ((p["u"], fu[r0:r0 + k], "u"), (p["v"], fv[r0:r0 + k], "v")):
            assert float(np.abs(a - b).max()) <= tol * max(1.0, float(np.abs(b).max())), what
    assert mismatched <= 8, mismatched
    assert sum(int(p["markers"]) for p in parts) == int(st.n_markers)
    if kind == "basic-fill" and not converged:
        assert int(st.pcg_iterations) == 100 * steps          # the solve ran, at the reference's cap
    ref.close()


@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("no_p2p", [False, True])
def test_host_program_two_ranks_match_one(no_p2p, tmp_path):
    """bin/euler-gpu --ranks 2 (two processes of the host C program, one GPU each, file
    rendezvous for the NCCL id and the NVLink peer handles) against the single-process run in
    the same modes: identical count plane (hash), marker total, sub-step and iteration counts.
    (Written after round 1's GPU budget was spent: not yet run on hardware.)"""
    import json
    import subprocess
    exe = os.path.join(ROOT, "bin", "euler-gpu")
    assert os.path.exists(exe), "run `make host`"
    src = tmp_path / "waterfall.txt"
    src.write_bytes(shipped_text("waterfall"))
    common = [exe, "--headless", "--frames", "25", "--grid", "160x96", "--precon", "rb", "--markers", "fast"]
    one = subprocess.run(common + [str(src)], check=True, capture_output=True, text=True, cwd=ROOT, timeout=300).stdout
    want = json.loads(one.strip().splitlines()[-1])
    rdv = tmp_path / "rdv"
    rdv.mkdir()
    procs = [subprocess.Popen(common + ["--ranks", "2", "--rank", str(r), "--rendezvous", str(rdv)] +
                              (["--no-p2p"] if no_p2p else []) + [str(src)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT) for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert [p.returncode for p in procs] == [0, 0], [o[1] for o in outs]
    got = json.loads(outs[0][0].strip().splitlines()[-1])
    assert outs[1][0].strip() == ""                       # only rank 0 reports
    for key in ("fnv_count", "markers", "substeps", "rng_state", "solves"):
        assert got[key] == want[key], key
    assert abs(got["pcg_iterations"] - want["pcg_iterations"]) <= got["solves"]    # +-1 per solve at the tolerance


def _rainbow_worker(rank, nranks, uid, text, nx, ny, substeps, out_dir):
    sys.path.insert(0, ROOT)
    from euler_b200 import gpu as G
    scn = Scenario(text, nx, ny)
    row0, rows = G.slab_partition(ny, nranks, rank)
    g = G.EulerGpu.from_scenario(scn, precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST, device=rank,
                                 slab_row0=row0, slab_rows=rows, rainbow=1, max_iterations=0)
    g.comm_init(rank, nranks, uid)
    g.comm_p2p_import(_exchange_blobs(out_dir, rank, nranks, g.comm_p2p_export()))
    for _ in range(substeps):
        g.substep(g.calculate_timestep(0.1))
    np.savez(os.path.join(out_dir, "rb%d.npz" % rank), row0=row0, rows=rows, count=g.get(G.F_COUNT),
             cr=g.get(G.F_CR), cg=g.get(G.F_CG), cb=g.get(G.F_CB))
    g.close()


@pytest.mark.parametrize("nranks", [2, 4])
@pytest.mark.parametrize("name,nx,ny,substeps", [("waterfall", 160, 96, 60), ("block", 100, 64, 60)])
def test_rainbow_on_slabs(name, nx, ny, substeps, nranks, tmp_path):
    """--rainbow colour transport (main.c:187-201, 424-438, 859-863, 873-882, 292-294) on row slabs
    against the ORACLE's --rainbow run (itself pinned to the reference's), iteration cap 0 on both
    sides so that the velocities are bit-determined: the colours of every fluid cell — what
    draw_rows() reads (main.c:938) — and the count plane bit for bit, with fluid, source colours and
    newly wet cells crossing the slab boundaries."""
    if _gpu_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    import torch.multiprocessing as mp
    from euler_b200 import gpu as G
    from oracle.oracle import Oracle
    text = resample(shipped_text(name), nx - 2, ny - 2)
    uid = G.comm_unique_id()
    mp.spawn(_rainbow_worker, args=(nranks, uid, text, nx, ny, substeps, str(tmp_path)), nprocs=nranks, join=True)
    o = Oracle(nx, ny, text, rainbow=True)
    o.c.precon_mode = 1; o.c.quirk_marker_dt_leak = 0; o.c.max_iterations = 0
    for _ in range(substeps):
        o.substep(o.calculate_timestep(0.1))
    parts = [np.load(os.path.join(str(tmp_path), "rb%d.npz" % r)) for r in range(nranks)]

    def merged(field, dtype):
        out = np.zeros((ny, nx), dtype)
        for p in parts:
            r0, n = int(p["row0"]), int(p["rows"])
            out[r0:r0 + n] = p[field][r0:r0 + n]
        return out
    assert same_bits(merged("count", np.uint8), o.count)
    fl = o.count != 0
    assert fl.any()
    for f, ref in (("cr", o.cr), ("cg", o.cg), ("cb", o.cb)):
        assert same_bits(merged(f, np.float32)[fl], ref[fl]), f


def _mixed_worker(rank, nranks, uid, text, nx, ny, frames, out_dir, p2p):
    sys.path.insert(0, ROOT)
    from euler_b200 import gpu as G
    scn = Scenario(text, nx, ny)
    row0, rows = G.slab_partition(ny, nranks, rank)
    g = G.EulerGpu.from_scenario(scn, precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST, device=rank,
                                 slab_row0=row0, slab_rows=rows, pcg_dtype=G.PCG_FP32, max_iterations=400)
    g.comm_init(rank, nranks, uid)
    if p2p:
        g.comm_p2p_import(_exchange_blobs(out_dir, rank, nranks, g.comm_p2p_export()))
    subs = [g.step_frame() for _ in range(frames)]
    st = g.stats()
    np.savez(os.path.join(out_dir, "mx%d.npz" % rank), row0=row0, rows=rows, count=g.get(G.F_COUNT),
             u=g.get(G.F_U), v=g.get(G.F_V), p=g.get(G.F_P), subs=np.array(subs), iters=st.pcg_iterations,
             markers=np.uint64(st.n_markers), rng=np.uint64(st.rng_state))
    g.close()


@pytest.mark.parametrize("p2p", [True, False])
@pytest.mark.parametrize("nranks,name,nx,ny,frames", [(2, "block", 100, 40, 12), (2, "waterfall", 160, 96, 20),
                                                      (4, "waterfall", 160, 96, 20)])
def test_mixed_precision_on_slabs(nranks, name, nx, ny, frames, p2p, tmp_path):
    """pcg_dtype = FP32 (fp32 PCG vectors, fp64 pressure and dot products, residual replacement:
    an opt-in mode that is not in the reference) on row slabs against the same mode on one GPU:
    classification bit-exact, marker totals, RNG state and sub-step counts equal, u, v and p within
    1e-5 (converged solves; the two runs differ in the summation order of the dot products only)."""
    if _gpu_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    import torch.multiprocessing as mp
    from euler_b200 import gpu as G
    text = shipped_text(name)
    if (nx, ny) != (100, 40):
        text = resample(text, nx - 2, ny - 2)
    uid = G.comm_unique_id()
    mp.spawn(_mixed_worker, args=(nranks, uid, text, nx, ny, frames, str(tmp_path), p2p), nprocs=nranks, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "mx%d.npz" % r)) for r in range(nranks)]
    ref = G.EulerGpu.from_scenario(Scenario(text, nx, ny), precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST,
                                   pcg_dtype=G.PCG_FP32, max_iterations=400)
    subs = [ref.step_frame() for _ in range(frames)]

    def merged(field, dtype):
        out = np.zeros((ny, nx), dtype)
        for p in parts:
            r0, n = int(p["row0"]), int(p["rows"])
            out[r0:r0 + n] = p[field][r0:r0 + n]
        return out
    st = ref.stats()
    assert all(list(p["subs"]) == subs for p in parts)
    assert same_bits(merged("count", np.uint8), ref.get(G.F_COUNT))
    assert sum(int(p["markers"]) for p in parts) == int(st.n_markers)
    assert all(int(p["rng"]) == int(st.rng_state) for p in parts)
    assert abs(int(parts[0]["iters"]) - int(st.pcg_iterations)) <= max(4, int(0.02 * st.pcg_iterations))
    for f, fld in (("u", G.F_U), ("v", G.F_V), ("p", G.F_P)):
        a, b = merged(f, np.float64 if f == "p" else np.float32), ref.get(fld)
        assert float(np.abs(a - b).max()) <= 1e-5 * max(1.0, float(np.abs(b).max())), f
    ref.close()
