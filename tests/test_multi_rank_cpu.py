"""world_size-2 gloo test (CPU) of the host-side multi-rank logic: the balanced row-slab
partition every rank computes for itself must tile the grid, and the rank plumbing bench.py
uses (broadcast of the 128-byte communicator id, max-over-ranks of the timing) must work."""
import os
import sys

import pytest

from conftest import ROOT


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from euler_b200 import gpu as G
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    res = []
    for ny in (40, 41, 1024, 16384, 17):
        r0, n = G.slab_partition(ny, world, rank)
        t = torch.tensor([r0, n], dtype=torch.int64)
        allt = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allt, t)
        rows = [(int(a[0]), int(a[1])) for a in allt]
        res.append((ny, rows))
    # communicator id: rank 0 makes 128 bytes, everyone ends up with the same bytes
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(uid, src=0)
    # timing: max over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        import json
        with open(out, "w") as f:
            json.dump({"parts": res, "uid_ok": bool((uid == torch.arange(128, dtype=torch.uint8)).all()),
                       "tmax": float(t[0])}, f)
    else:
        assert bool((uid == torch.arange(128, dtype=torch.uint8)).all())
    dist.destroy_process_group()


def test_slab_partition_and_rank_plumbing_gloo(tmp_path):
    import json
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "res.json")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = json.load(open(out))
    assert res["uid_ok"] and res["tmax"] == 11.0
    for ny, rows in res["parts"]:
        assert rows[0][0] == 0 and rows[0][0] + rows[0][1] == rows[1][0]
        assert rows[1][0] + rows[1][1] == ny and abs(rows[0][1] - rows[1][1]) <= 1


def test_slab_partition_many_ranks():
    from euler_b200 import gpu as G
    for ny in (16384, 8192, 1000, 37):
        for n in (1, 2, 4, 8):
            nxt = 0
            sizes = []
            for r in range(n):
                r0, k = G.slab_partition(ny, n, r)
                assert r0 == nxt
                nxt = r0 + k
                sizes.append(k)
            assert nxt == ny and max(sizes) - min(sizes) <= 1
    with pytest.raises(G.EulerGpuError):
        G.slab_partition(10, 2, 2)


def test_weighted_slab_partition():
    import numpy as np
    from euler_b200 import gpu as G
    ny = 4096
    w = np.ones(ny, np.uint64)
    w[: ny // 2] = 101                       # all the work in the lower half
    for n in (2, 4, 8):
        cuts, loads = [], []
        nxt = 0
        for r in range(n):
            r0, k = G.slab_partition_weighted(w, n, r)
            assert r0 == nxt and k >= 8
            nxt = r0 + k
            loads.append(int(w[r0:r0 + k].sum()))
        assert nxt == ny
        assert max(loads) <= 1.05 * (int(w.sum()) / n) + 101
    # degenerate: no weight anywhere still tiles the grid with >= 8 rows per slab
    z = np.zeros(64, np.uint64)
    rows = [G.slab_partition_weighted(z, 4, r) for r in range(4)]
    assert rows[0][0] == 0 and sum(k for _, k in rows) == 64 and all(k >= 8 for _, k in rows)


def test_weighted_slab_partition_properties():
    """Any weights (zeros, spikes, all the work in one row): the slabs tile [0, ny) in rank order
    and every slab keeps at least 2 x halo = 8 rows, which create() requires."""
    import numpy as np
    from hypothesis import given, settings, strategies as st
    from euler_b200 import gpu as G

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 8), st.integers(0, 300), st.integers(0, 2 ** 32 - 1), st.sampled_from(["flat", "spike", "zeros", "random"]))
    def run(n, extra, seed, kind):
        ny = 8 * n + extra
        rng = np.random.default_rng(seed)
        if kind == "flat":
            w = np.ones(ny, np.uint64)
        elif kind == "zeros":
            w = np.zeros(ny, np.uint64)
        elif kind == "spike":
            w = np.zeros(ny, np.uint64); w[int(rng.integers(0, ny))] = 10 ** 12
        else:
            w = rng.integers(0, 10 ** 6, ny).astype(np.uint64)
        nxt = 0
        for r in range(n):
            r0, k = G.slab_partition_weighted(w, n, r)
            assert r0 == nxt and k >= 8
            nxt = r0 + k
        assert nxt == ny

    run()
    with pytest.raises(G.EulerGpuError):
        G.slab_partition_weighted(np.ones(15, np.uint64), 2, 0)       # too short for two slabs
