#!/usr/bin/env python
"""tests/golden/make_golden.py — regenerate the committed fixtures from the REFERENCE.

Run in the build container (needs /root/reference and oracle/_ref built by
`make -C oracle ref`).  Writes, next to this file:

  scenarios.json       the five scenario texts the reference ships (input fixtures: the GPU
                       box has no /root/reference), keyed by name
  known_answers.json   known-answer vectors produced by the UNMODIFIED reference compiled with
                       the strict flags (oracle/build_ref.sh): for every scenario, at frames
                       0/1/10/50, marker count, fluid-cell count, FNV-1a of the uint8
                       marker-count plane, sum|u|, sum|v|, the RNG state and the number of
                       markers; plus a 64x48 resampled block scenario.  BASELINE.md §3 lists
                       the same quantities from the survey's probe.
  rainbow_answers.json known answers of the --rainbow colour transport (main.c:187-201, 424-438,
                       859-863, 873-882, 292-294) from the same reference build with
                       g_rainbow_enabled set: FNV-1a of the g_r, g_g, g_b planes masked to the
                       fluid cells (the only ones the renderer reads) at frames 0/1/10/30
  state_<name>_f<N>.npz   full reference state (u, v, counts, markers, precon) after N frames,
                       used as the common starting state of the per-stage parity tests.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle import Reference, fnv1a  # noqa: E402
from euler_b200.scenario import resample  # noqa: E402

REF_SCN = "/root/reference/scenarios"
NAMES = ["basic", "block", "filter", "waterfall", "weird-edges"]
FRAMES = [0, 1, 10, 50]


def snapshot(r):
    return {
        "markers": r.n_markers,
        "fluid_cells": int((r.count != 0).sum()),
        "fnv_count": "%016x" % fnv1a(r.count),
        "sum_abs_u": float(np.abs(r.u.astype(np.float64)).sum()),
        "sum_abs_v": float(np.abs(r.v.astype(np.float64)).sum()),
        "rng_state": "%016x" % r.rng_state,
    }


def rainbow_snapshot(r):
    fl = r.count != 0
    out = {"fluid_cells": int(fl.sum()), "fnv_count": "%016x" % fnv1a(r.count)}
    for k, plane in (("r", r.cr), ("g", r.cg), ("b", r.cb)):
        masked = np.where(fl, plane, np.float32(0)).astype(np.float32)
        out["fnv_" + k] = "%016x" % fnv1a(masked.view(np.uint8))
    return out


def main():
    scenarios = {}
    for n in NAMES:
        with open(os.path.join(REF_SCN, n + ".txt"), "rb") as f:
            scenarios[n] = f.read().decode("ascii")
    with open(os.path.join(HERE, "scenarios.json"), "w") as f:
        json.dump(scenarios, f, indent=0)

    answers = {}
    for n in NAMES:
        r = Reference(100, 40)
        r.init_from_text(scenarios[n])
        frame = 0
        for target in FRAMES:
            while frame < target:
                r.step_frame()
                frame += 1
            answers["%s@100x40/f%d" % (n, target)] = snapshot(r)
            if target == 10 and n in ("block", "waterfall", "weird-edges"):
                np.savez_compressed(os.path.join(HERE, "state_%s_f10.npz" % n), u=r.u, v=r.v,
                                    count=r.count, prev_count=r.prev_count, markers=r.markers.copy(),
                                    precon=r.precon, rng_state=np.uint64(r.rng_state),
                                    exhausted=np.uint8(r.source_exhausted))
    # a resampled (non-shipped size) case
    text = resample(scenarios["block"], 62, 46)
    r = Reference(64, 48)
    r.init_from_text(text)
    frame = 0
    for target in (0, 5, 20):
        while frame < target:
            r.step_frame()
            frame += 1
        answers["block@64x48/f%d" % target] = snapshot(r)
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(answers, f, indent=1, sort_keys=True)

    rainbow = {}
    for n in NAMES:
        r = Reference(100, 40)
        r.init_from_text(scenarios[n], rainbow=True)
        frame = 0
        for target in (0, 1, 10, 30):
            while frame < target:
                r.step_frame()
                frame += 1
            rainbow["%s@100x40/f%d" % (n, target)] = rainbow_snapshot(r)
    with open(os.path.join(HERE, "rainbow_answers.json"), "w") as f:
        json.dump(rainbow, f, indent=1, sort_keys=True)
    print("wrote %d known answers" % len(answers))


if __name__ == "__main__":
    main()
