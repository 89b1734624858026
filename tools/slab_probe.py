#!/usr/bin/env python
"""What does a PCG iteration cost on a THIN grid (the shape of one rank's slab at N=8) without
any exchange?  Single GPU, basic-fill NX x NY; per-iteration time = (sub-step at 2K iterations -
sub-step at K iterations) / K, no per-launch timers; then the per-kernel breakdown with timers.
Knobs are read from the environment by the library (one process per setting).

    python tools/slab_probe.py [NX] [NY] [K]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from euler_b200 import Scenario, synthetic
from euler_b200 import gpu as G


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 2150
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    scn = Scenario(synthetic(os.environ.get("PROBE_SCENARIO", "basic-fill"), nx, ny), nx, ny, row_major_markers=True)
    stream = torch.cuda.Stream()
    res = {}
    for iters in (k, 2 * k):
        sim = G.EulerGpu.from_scenario(scn, precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST,
                                       pcg_check_every=25, max_iterations=iters, stream=stream.cuda_stream)
        for _ in range(2):
            sim.substep(sim.calculate_timestep(0.1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        n = 4
        for _ in range(n):
            sim.substep(sim.calculate_timestep(0.1))
        e1.record(stream)
        torch.cuda.synchronize()
        res[iters] = e0.elapsed_time(e1) / n
        st = sim.stats()
        if iters == k:
            sim.set_profiling(True); sim.reset_profile()
            for _ in range(2):
                sim.substep(sim.calculate_timestep(0.1))
            prof = sim.kernel_profile()
            active = st.active_cells
        sim.close()
    per_it = (res[2 * k] - res[k]) / k
    print("grid %dx%d  active cells %d  env %s" % (nx, ny, active, {e: os.environ[e] for e in os.environ if e.startswith("EULER_")}))
    print("  sub-step %.3f ms at %d iterations, %.3f ms at %d  ->  %.4f ms per iteration (no timers), rest of the sub-step %.3f ms"
          % (res[k], k, res[2 * k], 2 * k, per_it, res[k] - k * per_it))
    alg = {"axpy_norm": 40, "rb_forward": 25, "rb_backward": 33, "fused_search_apply_a": 34, "fused_tail": 57}
    if "fused_tail" in prof:
        for k in ("axpy_norm", "rb_forward", "rb_backward"):
            alg.pop(k)
    tot = 0.0
    for name, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        avg = ms / cnt
        gbs = alg[name] * active / avg / 1e6 if name in alg else 0
        if name in alg:
            tot += avg
        print("  %-22s n %5d  avg %8.4f ms  %7.0f GB/s alg" % (name, cnt, avg, gbs))
    print("  sum of the iteration kernels with timers: %.4f ms; ideal at 6547 GB/s: %.4f ms" % (tot, sum(alg.values()) * active / 6547e6))


if __name__ == "__main__":
    main()
