#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 fluid-solve path (contract: see DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl gpu|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): MAC cell-updates/s = grid cells x sub-steps / time, whole job, plus PCG
iterations/s and the HBM-roofline fraction of the dominant kernel.  One "step" is one sub-step
of reference sim_step() (main.c:851-894): calculate_timestep, marker advection + re-binning,
sources, extrapolation, semi-Lagrangian velocity advection + gravity + boundaries, and the
pressure projection (PCG capped at the reference's 100 iterations).

Workload: the 16384^2 synthetic "basic-fill" scenario (walled box, fluid block
resting on the floor so the solve is active from the first sub-step), red-black IC(0)
preconditioner, fp64 PCG vectors as in the reference.  At N>1 the SAME grid is cut into N row
slabs, one per GPU/process (strong scaling): NCCL halo exchange, cross-slab marker migration,
PCG scalars reduced across ranks.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, built unmodified
from /root/reference with its own -O3 -ffast-math flags) on the box's host cores: one thread,
because the reference is single-threaded, on a bounded sample of the same workload (same
synthetic scenario at 1024^2; the metric is per cell, and at 16384^2 one CPU sub-step would
take ~13 minutes).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per cell of each kernel class, dense accounting over all cells with the
# reference's dtypes (fields fp32, PCG vectors fp64, masks u8): SURVEY §8d / DESIGN.md §4
ALG_BYTES_PER_CELL = {
    "apply_a": 18.0,            # R s 8 + fluid,a_diag 2 + W z 8
    "axpy_norm": 40.0,          # fused iteration: odd launches R As,r 16 + W r 8 = 24, even launches
                                # R As,r,p,s',s 40 + W r,p 16 = 56 (p takes two updates at once): mean 40;
                                # the reference's per-iteration form is 48 (R s,p,z,r 32 + W p,r 16)
    "precon_apply": 56.0,       # IC(0) wavefront: fwd R r,pc 16 W q 8; bwd R q,pc 16 W z 8; R r 8
    "rb_forward": 25.0,         # R r,pc 16 + fluid 1 + W q 8
    "rb_backward": 33.0,        # R q,pc,r 24 + fluid 1 + W z 8 (fused with z.r)
    "fused_search_apply_a": 34.0,  # R z,s 16 + fluid,a_diag 2 + W s',A s' 16
    "fused_axpy_forward": 65.0,    # R s,As,p,r,pc 40 + fluid 1 + W p,r',q 24
    "fused_tail": 57.0,            # axpy + forward + backward in one kernel (pcg_tail.cuh): odd launches
                                   # R r,As,pc 24 + fluid 1 + W r',z 16 = 41, even ones + R s',s,p 24 + W p 8 = 73
    "update_search": 24.0,      # R z,s 16 + W s 8
    "build_rhs": 27.0,          # R utmp,vtmp 8 + fluid,solid 2, W b 8 + a_diag 1 + p=0 8 (main.c:739)
    "pressure_update": 26.0,
    "extrapolate_bounds": 19.0,
    "advect_velocity": 18.0,
    "maxsq": 8.0,
}
# The grid-stage kernels leave a quad at once when its neighbourhood holds no fluid: there they
# move only the fluid mask and the zeros they must still write.  bytes per cell in that regime,
# for the "touched" figure reported next to the dense one (the dense accounting over all cells
# is SURVEY 8d's definition and stays the headline `frac` of those kernels).
DRY_BYTES_PER_CELL = {"build_rhs": 17.0, "pressure_update": 9.0, "extrapolate_bounds": 9.0,
                      "advect_velocity": 9.0}
ALG_BYTES_PER_MARKER = {"advect_markers": 16.0}
PCG_KERNELS = ("apply_a", "axpy_norm", "precon_apply", "update_search", "rb_forward", "rb_backward",
               "fused_search_apply_a", "fused_axpy_forward", "true_residual", "fused_tail")
# --pcg-dtype fp32 (euler_params.pcg_dtype = FP32, not the headline configuration): r, z, s, q,
# A s and the preconditioner diagonal are fp32 planes, p stays fp64 — DESIGN.md §9 row 4
ALG_BYTES_PER_CELL_FP32 = {
    "fused_search_apply_a": 18.0,  # R z,s 8 + fluid,a_diag 2 + W s',A s' 8
    "axpy_norm": 25.0,             # odd: R As,r 8 + fluid 1 + W r 4 = 13; even: + R s',s 8, p 8, W p 8 = 37
    "rb_forward": 13.0,            # R r,pc 8 + fluid 1 + W q 4
    "rb_backward": 17.0,           # R q,pc,r 12 + fluid 1 + W z 4
    "true_residual": 22.0,         # R p 8, b 8, fluid,a_diag 2 + W r 4; once every 10 iterations
}


NOMINAL_HBM_GBS = 8000.0     # north_star's "~8 TB/s" (DGX B200 figure); reported beside the measured peak


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def roofline_report(prof, alg_bytes, active_cells, cells_local, n_markers, headline, mixed,
                    grid_cells=None, wet_cells=None):
    """Per-kernel achieved GB/s of ALGORITHMIC bytes and the `roofline` object of the dominant
    kernel.  `prof` = {kernel class: (summed ms, launches)} from the library's CUDA-event timers.
    Units per launch: PCG kernels stream the tiles that contain fluid (like the reference, which
    touches only is_fluid cells): `active_cells`; marker kernels every marker; the grid stages the
    tiles of their list (`grid_cells`: tiles with fluid in or next to them within three sub-steps),
    where a quad without fluid in its neighbourhood moves only the fluid mask and the zeros it
    writes — so their bytes are B/cell x wet cells + (mask + zero store) x the other streamed
    cells, what the kernel really has to move (`frac`); the dense figure SURVEY 8d defines
    (B/cell x every stored cell, moved or not) is kept as `frac_dense` for reference only.
    `headline`: the workload the committed ncu capture (profiles/ncu_traffic.json) was taken on."""
    peak, peak_src = peaks()
    total_ms = sum(v[0] for v in prof.values())
    dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else None
    roof = None
    kernels = {}
    grid_cells = cells_local if grid_cells is None else grid_cells
    wet = min(active_cells if wet_cells is None else wet_cells, grid_cells)
    units_of = {}
    for name, (kms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        b_dense = None
        if name in PCG_KERNELS and name in alg_bytes:
            b = alg_bytes[name] * active_cells
            units_of[name] = active_cells
        elif name in DRY_BYTES_PER_CELL:
            b = alg_bytes[name] * wet + DRY_BYTES_PER_CELL[name] * (grid_cells - wet)
            b_dense = alg_bytes[name] * cells_local
            units_of[name] = grid_cells
        elif name in alg_bytes:
            b = alg_bytes[name] * cells_local
            units_of[name] = cells_local
        elif name in ALG_BYTES_PER_MARKER:
            b = ALG_BYTES_PER_MARKER[name] * n_markers
            units_of[name] = n_markers
        else:
            b = None
        avg = kms / cnt
        kernels[name] = {"ms_avg": round(avg, 4), "launches": cnt, "share": round(kms / total_ms, 4),
                         "gbs": round(b / avg / 1e6, 1) if b else None,
                         "frac": round(b / avg / 1e6 / peak, 4) if b else None,
                         "frac_nominal": round(b / avg / 1e6 / NOMINAL_HBM_GBS, 4) if b else None}
        if b_dense:
            kernels[name]["frac_dense"] = round(b_dense / avg / 1e6 / peak, 4)
            kernels[name]["bytes_per_launch"] = b
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tt = json.load(f)
        if headline and dom:
            traffic = (tt.get("fp32", {}) if mixed else tt).get(dom[0])
    except (OSError, ValueError):
        pass
    if dom and kernels[dom[0]]["gbs"]:
        k = kernels[dom[0]]
        units = units_of[dom[0]]
        roof = {"kernel": dom[0], "bound": "hbm", "achieved": k["gbs"], "peak": peak, "unit": "GB/s",
                "frac": k["frac"], "traffic": traffic, "peak_source": peak_src,
                "peak_nominal": NOMINAL_HBM_GBS, "frac_nominal": k["frac_nominal"],
                "traffic_source": "ncu capture committed under profiles/ (same workload)" if traffic else None,
                "alg_bytes_per_launch": alg_bytes[dom[0]] * units,
                "units_per_launch": units,
                "bytes_per_unit": alg_bytes[dom[0]],
                "ms_per_launch": k["ms_avg"], "share_of_step": k["share"]}
    return kernels, roof


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# Workloads: BASELINE.json `configs` plus the worst-case-traffic full-fluid box.
# name -> (scenario, grid, what it is)
CONFIGS = {
    "c5": ("basic-fill", 16384, "BASELINE config 5: 16384^2 synthetic basic-fill (the headline workload)"),
    "c3": ("waterfall", 4096, "BASELINE config 3: waterfall source/sink scenario resampled to 4096^2"),
    "c4": ("weird-edges", 8192, "BASELINE config 4: weird-edges irregular solid mask resampled to 8192^2"),
    "full": ("full", 16384, "full-fluid interior at 16384^2 (SURVEY 8d: worst-case traffic, every tile active)"),
}


def scenario_text(kind, nx, ny):
    """Scenario text at nx x ny: the two synthetic generators, or a shipped scenario file
    resampled nearest-neighbour, in the reference's scenario-file format."""
    from euler_b200 import synthetic, shipped_text, resample
    if kind in ("basic-fill", "full"):
        return synthetic(kind, nx, ny)
    return resample(shipped_text(kind), nx - 2, ny - 2)


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------- GPU arm ----

def run_gpu(args):
    import torch
    import torch.distributed as dist
    from euler_b200 import Scenario
    from euler_b200 import gpu as G

    rank, world, local = dist_env()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    n = args.grid
    precon = G.PRECON_REDBLACK if args.precon == "rb" else G.PRECON_IC0_WAVEFRONT
    mixed = args.pcg_dtype == "fp32"
    if mixed and args.precon != "rb":
        raise SystemExit("--pcg-dtype fp32 is a mode of the red-black solve")
    alg_bytes = dict(ALG_BYTES_PER_CELL, **ALG_BYTES_PER_CELL_FP32) if mixed else ALG_BYTES_PER_CELL

    t_host0 = time.perf_counter()
    text = scenario_text(args.scenario, n, n)
    # FAST marker mode: array order is free, store the seeded markers in row-major cell order
    scn = Scenario(text, n, n, row_major_markers=True)
    del text
    t_host = time.perf_counter() - t_host0

    stream = torch.cuda.Stream()          # the handle enqueues on this stream; events are recorded on it

    if world > 1 and args.precon != "rb":
        raise SystemExit("row slabs need --precon rb")
    # slabs balanced by work: the PCG and (since the tile list, common.cuh GridTiles) the grid stages
    # stream only tiles with fluid in or next to them, markers live in fluid cells; a dry row costs
    # one pass over its count bytes
    weight = scn.fluid.sum(axis=1, dtype=np.uint64) * 4096 + np.uint64(max(1, n // 256))
    row0, rows = G.slab_partition_weighted(weight, world, rank) if world > 1 else (0, 0)

    def make():
        sim = G.EulerGpu.from_scenario(scn, precon=precon, marker_mode=G.MARKERS_FAST,
                                       device=local, stream=stream.cuda_stream,
                                       pcg_check_every=args.check_every,
                                       slab_row0=row0, slab_rows=rows,
                                       pcg_dtype=G.PCG_FP32 if mixed else G.PCG_FP64)
        if world > 1:
            # communicator id made on rank 0, broadcast over torch.distributed (plumbing only)
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid = torch.tensor(list(G.comm_unique_id()), dtype=torch.uint8, device="cuda")
            dist.broadcast(uid, src=0)
            sim.comm_init(rank, world, bytes(uid.cpu().tolist()))
            if not args.no_p2p:
                # NVLink fast path: all-gather the CUDA-IPC blobs (plumbing), then map the peers
                mine = torch.tensor(list(sim.comm_p2p_export()), dtype=torch.uint8, device="cuda")
                allb = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(allb, mine)
                sim.comm_p2p_import([bytes(b.cpu().tolist()) for b in allb])
        return sim

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(sim):
        dt = sim.calculate_timestep(0.1)
        sim.substep(dt)

    # ---- device-resident timing (value) --------------------------------------------
    sim = make()
    cells = n * n
    # rows this rank stores (owned + halo rows): what the grid-stage kernels stream
    cells_local = cells if world == 1 else n * (rows + 8)
    for _ in range(args.warmup):
        one_step(sim)
    # (1) the timed region: exactly K steps, device time between two events on the handle's
    # stream, barrier + synchronize on both sides.  No per-launch timers in here: a CUDA event
    # pair around each of the ~420 launches of a step costs ~10 us of device time per pair (the
    # kernels before and after cannot overlap their tail/prologue across the timestamp) — 3 % of
    # a step on one GPU, 13 % on a thin slab of an 8-GPU run (profiles/r01e).
    sim.set_profiling(False)
    st0 = sim.stats()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        one_step(sim)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    st1 = sim.stats()
    # (2) the same K steps again, right away, with a CUDA event pair around every launch group
    # on the same stream: per-kernel durations for `roofline` and `kernels`
    prof, ms_timers = {}, None
    if not args.no_kernel_timers:
        sim.set_profiling(True)
        sim.reset_profile()
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(args.steps):
            one_step(sim)
        e3.record(stream)
        barrier()
        ms_timers = e2.elapsed_time(e3)
        prof = sim.kernel_profile()
        sim.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    iters = int(st1.pcg_iterations - st0.pcg_iterations)
    launches = int(st1.kernel_launches - st0.kernel_launches)
    n_markers = int(st1.n_markers)
    dev_bytes = int(st1.device_bytes)
    active_cells = int(st1.active_cells)
    grid_cells = int(st1.grid_cells)
    # ---- end to end through the C-ABI from host buffers (e2e) ------------------------
    # timed region: euler_gpu_reinit (== sim_init's hand-over: H2D of the three masks and the
    # seeded markers from pinned host arrays into the existing handle), K sub-steps, and after
    # every sub-step the D2H read of the marker-count plane into pinned host memory — what the
    # reference's step/draw loop moves (main.c:1034-1038).  Allocation and communicator
    # set-up are one-off and outside, like in the device-timed leg.
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    # Slab handles accept any superset of their own markers (include/euler_gpu.h): with the
    # markers in row-major cell order a slab's markers are one contiguous range of the global
    # array, so each rank ships ~1/N of it over PCIe.  The range is found from the marker rows
    # (host-side scenario preparation, like the parse itself; robust to any order: first / last
    # marker within one row of the slab, since init jitter can round a marker into the next row).
    markers_local = scn.markers
    if world > 1 and len(scn.markers):
        mrow = np.floor(scn.markers[:, 1]).astype(np.int32)
        inside = (mrow >= row0 - 1) & (mrow <= row0 + rows)
        if inside.any():
            lo_i = int(np.argmax(inside))
            hi_i = len(inside) - int(np.argmax(inside[::-1]))
            markers_local = scn.markers[lo_i:hi_i]
        else:
            markers_local = scn.markers[:0]
        del mrow, inside
    keep = [pinned(scn.solid), pinned(scn.source), pinned(scn.sink), pinned(markers_local)]
    (_, solid_h), (_, source_h), (_, sink_h), (_, markers_h) = keep
    count_host = torch.empty((n, n), dtype=torch.uint8, pin_memory=True).numpy()
    # what the reference's draw_rows() looks at (main.c:917-920) on a 240 x 67 terminal: rows
    # [max(Y-1-g_wy, 1), Y-1), columns [1, min(X-1, g_wx+1)) — the renderer feed of the host loop
    # (euler_gpu_read_window); --e2e-read plane reads the whole count plane instead
    wy0 = max(n - 1 - 67, 1)
    win = (1, wy0, min(n - 2, 240), n - 1 - wy0)
    barrier()
    t0 = time.perf_counter()
    sim.reinit(solid_h, source_h, sink_h, markers_h, scn.rng_state)
    t_reinit = time.perf_counter() - t0          # (euler_gpu_reinit returns after a stream synchronize)
    for _ in range(args.steps):
        one_step(sim)
        if args.e2e_read == "plane":
            sim.read_marker_count(count_host)
        else:
            sim.read_window(G.F_COUNT, win[0], win[1], win[2], win[3], count_host)
    sim.synchronize()
    t_e2e = time.perf_counter() - t0
    rows_stored = n if world == 1 else min(n, row0 + rows + 4) - max(0, row0 - 4)
    h2d_rank = 3 * n * rows_stored + markers_local.nbytes
    if args.e2e_read == "plane":
        d2h_rank = n * (n if world == 1 else rows)
    else:       # the part of the window this rank owns
        r_lo, r_hi = (0, n) if world == 1 else (row0, row0 + rows)
        d2h_rank = win[2] * max(0, min(r_hi, win[1] + win[3]) - max(r_lo, win[1]))
    # ---- invariants of the state K sub-steps after sim_init (outside every timed region) ----
    # the e2e leg started from euler_gpu_reinit, so this state does not depend on the warm-up:
    # runs on 1, 2, 4, 8 slabs with the same --steps must agree (integers exactly, sums to the
    # tolerance the unconverged solve allows — see DESIGN.md)
    chk = sim.check()
    st_end = sim.stats()
    migrated = int(st_end.markers_migrated)
    # ... and of a run whose every result is bit-determined: the same K sub-steps from sim_init with
    # the PCG iteration switched off (rhs, p = 0, velocity update only — marker advection and
    # hand-over between slabs, re-binning, sources, extrapolation, velocity advection, gravity,
    # bounds all run).  Its integers AND its count hash must be equal at every N, exactly.
    sim.set_max_iterations(0)
    sim.reinit(solid_h, source_h, sink_h, markers_h, scn.rng_state)
    for _ in range(args.steps):
        one_step(sim)
    chk0 = sim.check()
    rng0 = int(sim.stats().rng_state)
    sim.close()
    del sim, keep

    t = torch.tensor([ms, t_e2e * 1e3, t_reinit * 1e3], dtype=torch.float64, device="cuda")
    it = torch.tensor([launches, h2d_rank, d2h_rank, active_cells], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
    h2d = float(it[1]) / args.steps
    d2h = float(it[2])
    active_cells_all = int(it[3])     # halo-row tiles are counted by both neighbours: slightly above the N=1 figure
    ms_max, e2e_ms_max, reinit_ms_max = float(t[0]), float(t[1]), float(t[2])
    ksum = torch.tensor([sum(v[0] for v in prof.values()) / args.steps], dtype=torch.float64, device="cuda")
    ksums = [torch.zeros_like(ksum) for _ in range(world)]
    if world > 1:
        dist.all_gather(ksums, ksum)
    else:
        ksums = [ksum]
    per_rank_kernel_ms = [round(float(k[0]), 3) for k in ksums]
    iters_all, launches_all = iters, int(it[0])      # one global solve: every rank counts the same iterations
    # check object: integer fields add up over ranks modulo 2^64 (int64 wrap-around add), sums add, maxima max
    def as_i64(v):
        return int(np.array([v], dtype=np.uint64).view(np.int64)[0])
    ci = torch.tensor([as_i64(getattr(chk, k)) for k in ("n_markers", "fluid_cells", "count_sum", "count_hash")]
                      + [migrated] + [as_i64(getattr(chk0, k)) for k in ("n_markers", "fluid_cells", "count_sum", "count_hash")],
                      dtype=torch.int64, device="cuda")
    cs0 = torch.tensor([chk0.sum_abs_u, chk0.sum_abs_v], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cs0, op=dist.ReduceOp.SUM)
    cs = torch.tensor([chk.sum_abs_u, chk.sum_abs_v, chk.sum_p], dtype=torch.float64, device="cuda")
    cmx = torch.tensor([chk.max_abs_div, chk.max_abs_u, chk.max_abs_v], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ci, op=dist.ReduceOp.SUM)
        dist.all_reduce(cs, op=dist.ReduceOp.SUM)
        dist.all_reduce(cmx, op=dist.ReduceOp.MAX)
    ci = ci.cpu().numpy().view(np.uint64)
    check = {"state": "%d sub-steps after sim_init (end of the e2e leg)" % args.steps,
             "n_markers": int(ci[0]), "fluid_cells": int(ci[1]), "count_sum": int(ci[2]),
             "count_hash": "%016x" % int(ci[3]),
             "sum_abs_u": float(cs[0]), "sum_abs_v": float(cs[1]), "sum_p": float(cs[2]),
             "max_abs_div": float(cmx[0]), "max_abs_u": float(cmx[1]), "max_abs_v": float(cmx[2]),
             "pcg_iterations_last_solve": int(st_end.last_iterations), "last_residual": float(st_end.last_residual),
             "rng_state": "%016x" % int(st_end.rng_state), "substeps": int(st_end.substeps),
             "markers_migrated_total": int(ci[4]),
             "note": "the headline workload ends every solve at the reference's 100-iteration cap, far from "
                     "converged: runs on different N sum their dot products in different orders, CG amplifies that, "
                     "and after K sub-steps a few markers sit in a neighbouring cell — n_markers, fluid_cells and "
                     "rng_state agree exactly, the sums to ~1e-5 relative, count_hash only between runs that happen "
                     "to stay bit-identical; `no_solve` is the part that must agree exactly",
             "no_solve": {"state": "%d sub-steps after sim_init with max_iterations = 0 (every stage but the PCG "
                                   "iteration: bit-determined at any N)" % args.steps,
                          "n_markers": int(ci[5]), "fluid_cells": int(ci[6]), "count_sum": int(ci[7]),
                          "count_hash": "%016x" % int(ci[8]), "sum_abs_u": float(cs0[0]), "sum_abs_v": float(cs0[1]),
                          "rng_state": "%016x" % rng0}}

    if rank == 0:
        kernels, roof = roofline_report(prof, alg_bytes, active_cells, cells_local, n_markers,
                                        headline=(n == 16384 and world == 1 and args.scenario == "basic-fill"),
                                        mixed=mixed, grid_cells=grid_cells, wet_cells=int(chk.fluid_cells))
        value = cells * args.steps / (ms_max * 1e-3)
        line = {
            "metric": "MAC cell-updates/s", "value": value, "unit": "cell-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps,
            "ms_per_step_with_kernel_timers": (ms_timers / args.steps) if ms_timers else None,
            "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32 storage, f64 pressure and dot products" if mixed else "f64",
            "data": "synthetic",
            "config": {"workload": "%s %dx%d (whole grid), one sub-step of sim_step per step, PCG cap 100 "
                                   "(reference main.c:735), %s preconditioner, %s"
                                   % (args.scenario, n, n, "red-black IC(0)" if args.precon == "rb" else "IC(0) wavefront",
                                      "fp32 PCG vectors + fp64 p, residual replacement every 10 iterations "
                                      "(NOT the reference's precision: opt-in mode)" if mixed else "fp64 PCG vectors"),
                       "preset": args.config if (args.scenario, n) == CONFIGS[args.config][:2] else None,
                       "grid": [n, n], "markers": n_markers, "active_cells": active_cells,
                       "active_fraction": round(active_cells_all / cells, 4),
                       "fluid_cells": check["fluid_cells"], "fluid_fraction": round(check["fluid_cells"] / cells, 4),
                       "grid_stage_cells_rank0": grid_cells,
                       "markers_migrated_per_step": round(check["markers_migrated_total"] / max(1, int(st_end.substeps)), 1),
                       "parallelism": "single GPU" if world == 1 else
                                      "%d row slabs balanced by fluid cells (rank 0: %d rows), NCCL halo exchange + marker migration, %s" % (world, rows, "NCCL per-iteration exchanges" if args.no_p2p else "per-iteration exchanges by NVLink peer stores (CUDA IPC)"),
                       "l2_policy": "inputs >> L2: every plane is %.0f MB..%.0f MB vs 126 MB L2, no flush needed"
                                    % (cells / 1e6, cells * 8 / 1e6),
                       "device_bytes": dev_bytes, "host_setup_s": round(t_host, 2)},
            "pcg_iters_per_s": iters_all / (ms_max * 1e-3),
            "pcg_iterations": iters_all,
            "gpu_launches": launches_all,
            "e2e": {"value": cells * args.steps / (e2e_ms_max * 1e-3), "unit": "cell-updates/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "reinit_ms": round(reinit_ms_max, 2), "steps_ms": round(e2e_ms_max - reinit_ms_max, 2),
                    "note": "euler_gpu_reinit (sim_init hand-over: H2D of masks + seeded markers from pinned host arrays) "
                            "+ K sub-steps + per-step D2H of " + ("the whole count plane" if args.e2e_read == "plane" else
                            "the window of the count plane that draw_rows() reads (main.c:917-920, 240x67 terminal: "
                            "euler_gpu_read_window, the host loop's renderer feed)") +
                            "; handle allocation/communicator set-up outside"},
            "roofline": roof,
            "kernels": kernels,
            "per_rank_kernel_ms_per_step": per_rank_kernel_ms,
            "check": check,
            "clocks": clocks,
        }
        if world == 1 and not args.no_tol_study:
            ts = tol_study(G, Scenario, [int(x) for x in args.tol_study.split(",") if x], local)
            line["iters_to_tol"] = {k: {m: v[m]["iters"] for m in v} for k, v in ts.items()}
            line["ms_to_tol"] = {k: {m: v[m]["ms"] for m in v} for k, v in ts.items()}
            line["tol_study"] = {"what": "first projection of basic-fill NxN, iteration cap raised until ||r||inf <= 1e-6f "
                                         "(main.c:736): ic0 = the reference's natural-order IC(0) (wavefront kernels), "
                                         "rb = red-black IC(0); wall-clock ms of the whole solve on this GPU",
                                 "detail": ts}
        if not args.no_cpu and world >= 1:
            line["cpu_baseline"] = cpu_baseline(args, seconds=args.cpu_seconds)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def tol_study(G, Scenario, sizes, device):
    """Time-to-tolerance of both preconditioners (PCG iters/s is not time-to-solution):
    the first projection of basic-fill (hydrostatic column) with the iteration cap raised until
    ||r||inf <= 1e-6f (main.c:736, 756), natural-order IC(0) by wavefront (the reference's algorithm)
    and red-black IC(0); iterations and wall-clock milliseconds of the solve on this GPU."""
    out = {}
    for n in sizes:
        scn = Scenario(scenario_text("basic-fill", n, n), n, n, row_major_markers=True)
        res = {}
        for name, precon, every in (("rb", G.PRECON_REDBLACK, 50), ("ic0", G.PRECON_IC0_WAVEFRONT, 16)):
            sim = G.EulerGpu.from_scenario(scn, precon=precon, marker_mode=G.MARKERS_FAST, device=device,
                                           max_iterations=100000, pcg_check_every=every)
            dt = sim.calculate_timestep(0.1)
            for st in (G.S_ADVECT_MARKERS, G.S_REFRESH_COUNTS, G.S_SOURCES, G.S_EXTRAPOLATE, G.S_ADVECT_VELOCITY):
                sim.run_stage(st, dt)
            t0 = time.perf_counter()
            sim.run_stage(G.S_PROJECT, dt)
            ms = (time.perf_counter() - t0) * 1e3
            stt = sim.stats()
            res[name] = {"iters": int(stt.last_iterations), "ms": round(ms, 2), "residual": float(stt.last_residual)}
            sim.close()
        out[str(n)] = res
    return out


# ------------------------------------------------------------------- reference arm ----

def cpu_sample(args, steps, warmup):
    """Times the UNMODIFIED reference (oracle/_ref, its own -O3 -ffast-math flags) on one host
    core, same synthetic scenario resampled to the sample grid."""
    from oracle.oracle import Reference, ref_available, Oracle
    n = args.cpu_grid
    text = scenario_text(args.scenario, n, n)
    if ref_available(n, n, fast=True):
        sim = Reference(n, n, fast=True)
        sim.init_from_text(text)
        kind = "reference"

        def step():
            # one sub-step == sim_step() with the frame cut after the first sub-step is not
            # expressible without touching the reference; time its own stages in sim_step order
            dt = sim.calculate_timestep(0.1)
            sim.advect_markers(dt); sim.refresh_marker_counts(); sim.update_fluid_sources()
            sim.extrapolate(sim.u, 1); sim.extrapolate(sim.v, 2)
            sim.zero_bounds(sim.u, 1); sim.zero_bounds(sim.v, 2)
            sim.advect_u(dt); sim.advect_v(dt); sim.apply_body_forces(dt)
            sim.zero_bounds(sim.utmp, 1); sim.zero_bounds(sim.vtmp, 2)
            sim.project(dt)
    else:
        sim = Oracle(n, n, text)
        kind = "port"

        def step():
            sim.substep(sim.calculate_timestep(0.1))
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": n * n * steps / dt, "unit": "cell-updates/s", "cores": 1, "kind": kind,
            "host_cores_available": os.cpu_count(),
            "sample": "%s resampled to %dx%d, %d sub-steps (PCG cap 100), single thread: the reference "
                      "is single-threaded" % (args.scenario, n, n, steps),
            "seconds": dt, "ms_per_step": dt / steps * 1e3}


def cpu_baseline(args, seconds=20.0):
    # ~3 s per sub-step at 1024^2: 1 warm-up + enough steps for ~seconds
    steps = max(1, int(seconds / 3.5))
    return cpu_sample(args, steps, 1)


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    steps = max(1, min(args.steps, 6))
    warm = 1 if args.warmup > 0 else 0
    b = cpu_sample(args, steps, warm)
    line = {"impl": "reference", "metric": "MAC cell-updates/s", "value": b["value"],
            "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": b["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s, reference CPU path on a %dx%d sample of the %dx%d workload"
                                   % (args.scenario, args.cpu_grid, args.cpu_grid, args.grid, args.grid)},
            "cpu_baseline": b,
            "e2e": {"value": b["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS),
                    help="workload preset (BASELINE.json configs 3-5, or the full-fluid worst case); "
                         "--scenario / --grid override its parts")
    ap.add_argument("--grid", type=int, default=None)
    ap.add_argument("--scenario", default=None,
                    help="basic-fill | full (synthetic) or a shipped scenario name (resampled to --grid)")
    ap.add_argument("--precon", default="rb", choices=["rb", "ic0"])
    ap.add_argument("--check-every", type=int, default=25)
    ap.add_argument("--pcg-dtype", default="fp64", choices=["fp64", "fp32"],
                    help="fp32: opt-in mixed-precision PCG (not the headline configuration)")
    ap.add_argument("--cpu-grid", type=int, default=1024)
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--tol-study", default="1024,4096", help="grid sizes of the time-to-tolerance study (N=1 only)")
    ap.add_argument("--no-tol-study", action="store_true")
    ap.add_argument("--e2e-read", default="window", choices=["window", "plane"],
                    help="what the e2e leg reads back every step: the renderer's window of the count plane "
                         "(what the host loop does) or the whole plane")
    ap.add_argument("--no-kernel-timers", action="store_true",
                    help="no per-launch CUDA events inside the timed region (no roofline/kernels objects): "
                         "measures what the events themselves cost")
    ap.add_argument("--no-p2p", action="store_true", help="N>1: NCCL for every exchange (no CUDA-IPC fast path)")
    args = ap.parse_args()
    if args.scenario is None:
        args.scenario = CONFIGS[args.config][0]
    if args.grid is None:
        args.grid = CONFIGS[args.config][1]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
