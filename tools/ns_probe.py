#!/usr/bin/env python
"""Ring-depth A/B of the fp32 (pcg_dtype = FP32) pipe kernels: one scenario, one handle per
EULER_NS_MIXED value, per-kernel averages from the library's CUDA-event timers (development aid).

    python tools/ns_probe.py [N] [ns,ns,...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from euler_b200 import Scenario, synthetic
from euler_b200 import gpu as G


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    depths = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "4,6,8").split(",")]
    t0 = time.time()
    scn = Scenario(synthetic("basic-fill", n, n), n, n, row_major_markers=True)
    print("host setup %.1fs" % (time.time() - t0), flush=True)
    names = ("fused_search_apply_a", "axpy_norm", "rb_forward", "rb_backward", "true_residual")
    for ns in depths:
        os.environ["EULER_NS_MIXED"] = str(ns)          # read once per handle at create()
        sim = G.EulerGpu.from_scenario(scn, precon=G.PRECON_REDBLACK, marker_mode=G.MARKERS_FAST,
                                       pcg_check_every=25, pcg_dtype=G.PCG_FP32)
        sim.substep(sim.calculate_timestep(0.1))
        sim.set_profiling(True); sim.reset_profile()
        for _ in range(2):
            sim.substep(sim.calculate_timestep(0.1))
        prof = sim.kernel_profile()
        st = sim.stats()
        per_it = sum(prof[k][0] for k in names if k in prof) / max(1, prof["axpy_norm"][1])
        print("NS=%d  " % ns + "  ".join("%s %.4f" % (k.split("_")[0] + k[-2:], prof[k][0] / prof[k][1]) for k in names if k in prof)
              + "  | ms/iteration %.4f  resid %.3e" % (per_it, st.last_residual), flush=True)
        sim.close()


if __name__ == "__main__":
    main()
