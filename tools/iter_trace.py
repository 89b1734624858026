#!/usr/bin/env python
"""Where does a red-black PCG iteration spend its time?  In-kernel timeline (EULER_TRACE,
include/euler_gpu.h euler_gpu_trace_read) of the two iteration kernels on 1..N row slabs:
per kernel the launch gap, the ramp of block starts, the wait for the other ranks' scalars,
the row loop, the spread of block ends and the epilogue (fence + reduction + post).

    python tools/iter_trace.py [NX] [NY]                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 \
        tools/iter_trace.py [NX] [NY]                                      # N row slabs
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from euler_b200 import Scenario, synthetic
from euler_b200 import gpu as G


def analyse(tr):
    """tr: [n, 16] uint64 -> {kind: {phase: mean microseconds}} over the launches after the first few"""
    inv = lambda a: (~a).astype(np.int64)
    val = lambda a: a.astype(np.int64)
    out = {}
    prev_exit = None
    rows = {1: [], 2: []}
    for s in tr:
        kind = int(s[8])
        st0, st1, c0, c1, w0, w1, e0, e1 = inv(s[0]), val(s[1]), inv(s[2]), val(s[3]), inv(s[4]), val(s[5]), inv(s[6]), val(s[7])
        if kind in rows and prev_exit is not None and s[1] and s[7]:
            rows[kind].append([st0 - prev_exit, st1 - st0, c0 - st0, c1 - st1, w0 - c0, w1 - c1, w1 - w0, e1 - w1,
                               e1 - prev_exit, e1 - st0])
        prev_exit = e1 if s[7] else prev_exit
    names = ["gap", "ramp", "collect_first", "collect_last", "rows_first", "rows_last", "end_spread", "epilogue",
             "period", "kernel"]
    for kind, r in rows.items():
        if len(r) > 12:
            a = np.array(r[8:], dtype=np.float64) / 1e3
            out[kind] = dict(zip(names, a.mean(axis=0).round(2)), n=len(a),
                             period_p10=round(float(np.percentile(a[:, 8], 10)), 2),
                             period_p90=round(float(np.percentile(a[:, 8], 90)), 2))
    return out


def main():
    os.environ.setdefault("EULER_TRACE", "4096")      # read by the library when the handle is created
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    scn = Scenario(synthetic("basic-fill", nx, ny), nx, ny, row_major_markers=True)
    weight = scn.fluid.sum(axis=1, dtype=np.uint64) * 4096 + np.uint64(max(1, ny // 256))
    row0, rows = G.slab_partition_weighted(weight, world, rank) if world > 1 else (0, 0)
    stream = torch.cuda.Stream()
    sim = G.EulerGpu.from_scenario(scn, precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST, device=local,
                                   stream=stream.cuda_stream, pcg_check_every=25, slab_row0=row0, slab_rows=rows)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(G.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, src=0)
        sim.comm_init(rank, world, bytes(uid.cpu().tolist()))
        mine = torch.tensor(list(sim.comm_p2p_export()), dtype=torch.uint8, device="cuda")
        allb = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allb, mine)
        sim.comm_p2p_import([bytes(b.cpu().tolist()) for b in allb])
    for _ in range(3):
        sim.substep(sim.calculate_timestep(0.1))
    sim.trace_read()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.substep(sim.calculate_timestep(0.1))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tr = sim.trace_read()
    res = analyse(tr)
    lines = ["rank %d rows %d  sub-step %.3f ms  %d traced launches  env %s" % (
        rank, rows or ny, ms, len(tr), {k: v for k, v in os.environ.items() if k.startswith("EULER_") and k != "EULER_TRACE"})]
    for kind, name in ((1, "search+apply"), (2, "tail")):
        if kind in res:
            lines.append("  %-12s %s" % (name, " ".join("%s=%s" % kv for kv in res[kind].items())))
    tb = getattr(sim, "trace_blocks", None)
    if os.environ.get("EULER_TRACE_BLOCKS") and tb is not None:
        b = tb[tb[:, 0] > 0].astype(np.int64)
        if len(b):
            t0 = b[:, 0].min()
            st, en, sm = (b[:, 0] - t0) / 1e3, (b[:, 1] - t0) / 1e3, b[:, 2]
            dur = en - st
            q = lambda a: " ".join("%.1f" % v for v in np.percentile(a, [0, 10, 25, 50, 75, 90, 100]))
            lines.append("  blocks of kernel %s, last launch: n=%d  start[us] %s | end %s | duration %s" % (
                os.environ["EULER_TRACE_BLOCKS"], len(b), q(st), q(en), q(dur)))
            order = np.argsort(en)
            lines.append("    first done (block:sm:end) " + " ".join("%d:%d:%.1f" % (i, sm[i], en[i]) for i in order[:12]))
            lines.append("    last done  (block:sm:end) " + " ".join("%d:%d:%.1f" % (i, sm[i], en[i]) for i in order[-12:]))
            # by block index (position in the row split) and by SM
            nb = len(b); g = max(1, nb // 16)
            lines.append("    mean end by block-index group of %d: " % g + " ".join("%.1f" % en[i:i + g].mean() for i in range(0, nb, g)))
            per_sm = {}
            for i in range(nb):
                per_sm.setdefault(int(sm[i]), []).append(en[i])
            sm_mean = np.array([np.mean(v) for v in per_sm.values()])
            lines.append("    SMs %d, blocks per SM %s, mean end per SM: %s ; corr(end of the two slowest per SM) n/a" % (
                len(per_sm), sorted(set(len(v) for v in per_sm.values())), q(sm_mean)))
            same = [max(v) - min(v) for v in per_sm.values() if len(v) > 1]
            if same:
                lines.append("    spread of block ends WITHIN an SM: %s" % q(np.array(same)))
    text = "\n".join(lines)
    if world > 1:
        allt = [None] * world
        dist.all_gather_object(allt, text)
        if rank == 0:
            print("\n".join(allt))
        dist.destroy_process_group()
    else:
        print(text)
    print("(microseconds, means over the launches of one sub-step; gap = previous kernel's last block exit -> first "
          "block start; collect = start -> scalars known, first / last block; rows = the row loop; end_spread = "
          "first -> last block done; epilogue = last block done -> kernel exit)") if rank == 0 else None
    sim.close()


if __name__ == "__main__":
    main()
