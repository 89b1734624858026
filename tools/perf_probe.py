#!/usr/bin/env python
"""Per-kernel timing probe at a given grid size (development aid).

    python tools/perf_probe.py [N] [rb|ic0] [steps] [fp64|fp32]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from euler_b200 import Scenario, synthetic
from euler_b200 import gpu as G

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    precon = G.PRECON_REDBLACK if (len(sys.argv) < 3 or sys.argv[2] == "rb") else G.PRECON_IC0_WAVEFRONT
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    mixed = len(sys.argv) > 4 and sys.argv[4] == "fp32"
    t0 = time.time()
    scn = Scenario(synthetic("basic-fill", n, n), n, n, row_major_markers=os.environ.get("EULER_ROWMAJOR", "1") == "1")
    print("host setup %.1fs markers %d" % (time.time() - t0, len(scn.markers)))
    t0 = time.time()
    sim = G.EulerGpu.from_scenario(scn, precon=precon, marker_mode=G.MARKERS_FAST, pcg_check_every=25,
                                   stencil_variant=int(os.environ.get('EULER_VARIANT', '0')),
                                   pcg_dtype=G.PCG_FP32 if mixed else G.PCG_FP64)
    print("create %.2fs device GB %.2f" % (time.time() - t0, sim.stats().device_bytes / 1e9))
    sim.substep(sim.calculate_timestep(0.1))
    sim.set_profiling(True); sim.reset_profile()
    t0 = time.time()
    for _ in range(steps):
        sim.substep(sim.calculate_timestep(0.1))
    wall = time.time() - t0
    st = sim.stats()
    cells = n * n
    print("grid %d precon %d %s: %.1f ms/substep, iters %d resid %.3e" % (n, precon, "fp32" if mixed else "fp64", wall / steps * 1e3, st.last_iterations, st.last_residual))
    bpc = {"build_rhs": 19, "pressure_update": 26, "extrapolate_bounds": 19, "advect_velocity": 18, "maxsq": 8}
    pcg = {"apply_a": 18, "axpy_norm": 40, "precon_apply": 56, "update_search": 24, "rb_forward": 25,
           "rb_backward": 33, "fused_search_apply_a": 34, "fused_axpy_forward": 65}
    if mixed:
        pcg.update({"fused_search_apply_a": 18, "axpy_norm": 25, "rb_forward": 13, "rb_backward": 17, "true_residual": 22})
    active = st.active_cells
    for name, (ms, cnt) in sorted(sim.kernel_profile().items(), key=lambda kv: -kv[1][0]):
        avg = ms / cnt
        gbs = (bpc[name] * cells / avg / 1e6 if name in bpc else pcg[name] * active / avg / 1e6 if name in pcg
               else 16 * st.n_markers / avg / 1e6 if name == "advect_markers" else 0)
        print("  %-20s total %9.2f ms  n %5d  avg %8.4f ms  %7.0f GB/s alg" % (name, ms, cnt, avg, gbs))

if __name__ == "__main__":
    main()
