import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch.multiprocessing as mp
from euler_b200 import Scenario, shipped_text, resample

def exch(out_dir, rank, nranks, blob):
    with open(os.path.join(out_dir, "b%d.tmp" % rank), "wb") as f: f.write(blob)
    os.rename(os.path.join(out_dir, "b%d.tmp" % rank), os.path.join(out_dir, "b%d.bin" % rank))
    res = []
    for r in range(nranks):
        p = os.path.join(out_dir, "b%d.bin" % r)
        while not os.path.exists(p): time.sleep(0.01)
        res.append(open(p, "rb").read())
    return res

def worker(rank, nranks, uid, text, nx, ny, p2p, out_dir):
    from euler_b200 import gpu as G
    scn = Scenario(text, nx, ny)
    row0, rows = G.slab_partition(ny, nranks, rank)
    g = G.EulerGpu.from_scenario(scn, precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST, device=rank, slab_row0=row0, slab_rows=rows)
    g.comm_init(rank, nranks, uid)
    if p2p: g.comm_p2p_import(exch(out_dir, rank, nranks, g.comm_p2p_export()))
    log = open(os.path.join(out_dir, "log%d" % rank), "w")
    for f in range(14):
        n = g.step_frame(); st = g.stats()
        if f >= 9: log.write("p2p=%d rank %d frame %02d substeps %d iters %d resid %.6e markers %d total_iters %d\n" % (p2p, rank, f, n, st.last_iterations, st.last_residual, st.n_markers, st.pcg_iterations))
    log.close()
    g.close()

if __name__ == "__main__":
    import tempfile
    from euler_b200 import gpu as G
    text = shipped_text("block")
    for p2p in (0, 1):
        d = tempfile.mkdtemp()
        uid = G.comm_unique_id()
        mp.spawn(worker, args=(2, uid, text, 100, 40, p2p, d), nprocs=2, join=True)
        for r in range(2): print(open(os.path.join(d, "log%d" % r)).read())
