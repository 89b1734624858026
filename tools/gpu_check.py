#!/usr/bin/env python
"""Quick GPU-vs-oracle sanity run (development aid; the real checks live in tests/)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from euler_b200 import Scenario, shipped_text, resample
from euler_b200 import gpu as G
from oracle.oracle import Oracle

def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a.view(np.uint64) if a.dtype == np.float64 else a

def cmp(name, a, b, mask=None):
    if mask is not None:
        a = a[mask]; b = b[mask]
    same = np.array_equal(bits(np.ascontiguousarray(a)), bits(np.ascontiguousarray(b)))
    if a.dtype.kind == 'f':
        d = np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64))) if a.size else 0.0
        print("  %-22s bit-exact=%s maxabs=%.3e" % (name, same, d))
    else:
        print("  %-22s equal=%s ndiff=%d" % (name, same, int((a != b).sum())))
    return same

def load_state(o, g):
    g.set(G.F_U, o.u); g.set(G.F_V, o.v); g.set(G.F_UTMP, o.utmp); g.set(G.F_VTMP, o.vtmp)
    g.set(G.F_COUNT, o.count); g.set(G.F_PREV_COUNT, o.prev_count)
    g.set(G.F_MARKERS, o.markers); g.set(G.F_PRECON, o.precon)
    g.set_rng_state(int(o.c.rng_state)); g.set_source_exhausted(int(o.c.source_exhausted))

def run(name, nx, ny, frames, precon):
    print("== %s %dx%d precon=%d" % (name, nx, ny, precon))
    text = shipped_text(name)
    if (nx, ny) != (100, 40):
        text = resample(text, nx - 2, ny - 2)
    scn = Scenario(text, nx, ny)
    o = Oracle(nx, ny, text)
    o.c.quirk_marker_dt_leak = 0
    o.c.precon_mode = precon
    cmp("host markers", scn.markers, o.markers)
    assert scn.rng_state == int(o.c.rng_state)
    g = G.EulerGpu.from_scenario(scn, precon=precon, marker_mode=G.MARKERS_FAST)
    cmp("init count", g.get(G.F_COUNT), o.count)
    for f in range(frames):
        o.step_frame()
    load_state(o, g)
    fl = o.count != 0
    # one sub-step, stage by stage
    dt_o = o.calculate_timestep(0.1); dt_g = g.calculate_timestep(0.1)
    print("  dt oracle %.9g gpu %.9g" % (dt_o, dt_g))
    dt = dt_o
    o.advect_markers(dt); g.run_stage(G.S_ADVECT_MARKERS, dt)
    cmp("advect_markers", g.get(G.F_MARKERS), o.markers)
    o.refresh_marker_counts(); g.run_stage(G.S_REFRESH_COUNTS)
    cmp("refresh: count", g.get(G.F_COUNT), o.count)
    cmp("refresh: prev", g.get(G.F_PREV_COUNT), o.prev_count)
    cmp("refresh: markers", g.get(G.F_MARKERS), o.markers)
    o.update_fluid_sources(); g.run_stage(G.S_SOURCES)
    cmp("sources: count", g.get(G.F_COUNT), o.count)
    cmp("sources: markers", g.get(G.F_MARKERS), o.markers)
    print("  rng", hex(int(o.c.rng_state)), hex(int(g.stats().rng_state)))
    o.extrapolate(o.u, 1); o.extrapolate(o.v, 2); o.zero_bounds(o.u, 1); o.zero_bounds(o.v, 2)
    g.run_stage(G.S_EXTRAPOLATE)
    cmp("extrapolate u", g.get(G.F_U), o.u); cmp("extrapolate v", g.get(G.F_V), o.v)
    o.advect_u(dt); o.advect_v(dt); o.apply_body_forces(dt); o.zero_bounds(o.utmp, 1); o.zero_bounds(o.vtmp, 2)
    g.run_stage(G.S_ADVECT_VELOCITY, dt)
    cmp("advect utmp", g.get(G.F_UTMP), o.utmp); cmp("advect vtmp", g.get(G.F_VTMP), o.vtmp)
    fl = o.count != 0
    o.build_rhs(dt); g.run_stage(G.S_BUILD_RHS, dt)
    cmp("rhs b", g.get(G.F_R), o.b); cmp("adiag", g.get(G.F_ADIAG), o.adiag, fl)
    o.r[:] = o.b
    o.apply_preconditioner(o.r, o.z); g.run_stage(G.S_PRECONDITION)
    cmp("precon plane", g.get(G.F_PRECON), o.precon)
    cmp("q", g.get(G.F_Q), o.q, fl); cmp("z=M^-1 r", g.get(G.F_Z), o.z, fl)
    o.s[:] = o.z; g.set(G.F_S, o.z)
    o.apply_a(o.s, o.z); g.run_stage(G.S_APPLY_A)
    cmp("z=A s", g.get(G.F_Z), o.z, fl)
    # the full project from the same utmp/vtmp
    o.project(dt); g.run_stage(G.S_PROJECT, dt)
    st = g.stats()
    print("  iters oracle %d gpu %d  resid oracle %.3e gpu %.3e" % (o.c.last_iterations, st.last_iterations, o.c.last_residual, st.last_residual))
    cmp("p", g.get(G.F_P), o.p, fl)
    cmp("u", g.get(G.F_U), o.u); cmp("v", g.get(G.F_V), o.v)
    # whole frames
    t0 = time.time()
    for f in range(5):
        o.step_frame(); g.step_frame()
    print("  5 frames: %.2fs; launches %d" % (time.time() - t0, g.stats().kernel_launches))
    cmp("frames: count", g.get(G.F_COUNT), o.count)
    cmp("frames: u", g.get(G.F_U), o.u); cmp("frames: v", g.get(G.F_V), o.v)
    gm = g.get(G.F_MARKERS); print("  markers", len(gm), o.n_markers)
    if len(gm) == o.n_markers: cmp("frames: markers", gm, o.markers)
    g.close()

if __name__ == "__main__":
    run("block", 100, 40, 12, 0)
    run("waterfall", 100, 40, 30, 0)
    run("block", 100, 40, 12, 1)
    run("weird-edges", 256, 256, 6, 0)
    run("waterfall", 256, 256, 8, 1)
