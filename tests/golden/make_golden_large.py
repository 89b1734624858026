#!/usr/bin/env python
"""tests/golden/make_golden_large.py — digests of ORACLE runs at BASELINE's large configurations,
for the parity tests that run where a CPU oracle pass would be too slow or too expensive (an
8192^2 oracle sub-step costs minutes of CPU, and a multi-GPU box is charged per GPU).

Run in the build container (CPU only; ~10 GB of RAM for 8192^2):  python tests/golden/make_golden_large.py
Writes large_answers.json next to this file.

Cases (oracle = oracle/liboracle.so, the restatement pinned bit for bit to the unmodified reference
by tests/test_oracle.py; strict IEEE flags, quirk_marker_dt_leak = 0 == EULER_MARKERS_FAST):

  weird-edges 8192^2, max_iterations = 0, S sub-steps (BASELINE config 4 geometry).  With the
      iteration cap at 0 project() is rhs + p = 0 + the velocity update, so every stage except
      the PCG iteration runs (marker advection through the irregular solids, re-binning and
      deletion, extrapolation, velocity advection, gravity, bounds) and every result is
      determined bit for bit: the slab-decomposed GPU run must reproduce the count plane, the
      marker MULTISET and u, v exactly.  After each sub-step: FNV-1a of the count plane, of the
      marker array sorted by (y bits, x bits), of the u and v planes; marker total; RNG state.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, fnv1a  # noqa: E402
from euler_b200.scenario import resample, shipped_text  # noqa: E402


def sorted_marker_digest(m):
    """FNV-1a of the markers sorted by (y bits, x bits): the multiset, independent of array order."""
    b = np.ascontiguousarray(m, dtype=np.float32).view(np.uint32).reshape(-1, 2)
    key = (b[:, 1].astype(np.uint64) << np.uint64(32)) | b[:, 0].astype(np.uint64)
    key.sort()
    return "%016x" % fnv1a(key.view(np.uint8))


def digest(o):
    return {"markers": o.n_markers, "fluid_cells": int((o.count != 0).sum()),
            "fnv_count": "%016x" % fnv1a(o.count), "fnv_markers_sorted": sorted_marker_digest(o.markers),
            "fnv_u": "%016x" % fnv1a(o.u.view(np.uint8)), "fnv_v": "%016x" % fnv1a(o.v.view(np.uint8)),
            "rng_state": "%016x" % int(o.c.rng_state), "dt": float(o.c.last_dt)}


def no_solve_case(name, n, substeps):
    text = resample(shipped_text(name), n - 2, n - 2)
    o = Oracle(n, n, text)
    o.c.quirk_marker_dt_leak = 0
    o.c.precon_mode = 1
    o.c.max_iterations = 0
    out = {"scenario": name, "grid": [n, n], "max_iterations": 0, "substeps": []}
    for i in range(substeps):
        t0 = time.time()
        o.substep(o.calculate_timestep(0.1))
        out["substeps"].append(digest(o))
        print(name, n, "sub-step", i, out["substeps"][-1], "%.1fs" % (time.time() - t0), flush=True)
    return out


def main():
    cases = {"weird-edges_8192_nosolve": no_solve_case("weird-edges", 8192, 6),
             "weird-edges_512_nosolve": no_solve_case("weird-edges", 512, 6)}
    with open(os.path.join(HERE, "large_answers.json"), "w") as f:
        json.dump(cases, f, indent=1)


if __name__ == "__main__":
    main()
