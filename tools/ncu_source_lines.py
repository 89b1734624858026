#!/usr/bin/env python
"""Per-SOURCE-LINE cost of one kernel from an `ncu --set full --import-source on` report (the
library is compiled with -lineinfo): warp-level instructions executed and stall samples per
line of CUDA source, plus the opcode mix and the stall reasons of the whole kernel.

    ncu_source_lines.py <report.ncu-rep> <kernel regex> [launch index] [top N]      -> markdown on stdout"""
import collections
import csv
import re
import subprocess
import sys


def run(rep, kern, launch, view):
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", view, "--csv", "--kernel-name", "regex:" + kern,
           "--launch-skip", str(launch), "--launch-count", "1"]
    return subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 18
    # ---- SASS view: opcode mix + stall reasons
    lines = run(rep, kern, launch, "sass")
    start = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
    name = next(csv.reader([lines[start[0]]]))[1]
    rows = list(csv.reader(lines[start[0] + 1: start[1] if len(start) > 1 else None]))
    hdr = rows[0]
    rows = [r for r in rows[1:] if len(r) == len(hdr) and r != hdr]
    ix = {h: i for i, h in enumerate(hdr)}
    n_inst = sum(int(r[ix["Instructions Executed"]]) for r in rows)
    n_smp = sum(int(r[ix["# Samples"]]) for r in rows)
    mix, smp = collections.Counter(), collections.Counter()
    for r in rows:
        op = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip()).split()[0].split(".")[0]
        mix[op] += int(r[ix["Instructions Executed"]]); smp[op] += int(r[ix["# Samples"]])
    stalls = collections.Counter({h: sum(int(r[ix[h]]) for r in rows) for h in hdr
                                 if h.startswith("stall_") and "Not Issued" not in h})
    short = re.sub(r"\(euler::Grid.*", "", name).replace("euler::<unnamed>::", "").replace("(int)", "").replace("void ", "")
    print("## %s (launch %d of the capture)\n" % (short, launch))
    print("%d warp-level instructions, %d static SASS instructions, %d stall samples\n" % (n_inst, len(rows), n_smp))
    print("opcode mix (executed / samples): " + ", ".join(
        "%s %.1f%% / %.1f%%" % (op, 100.0 * n / n_inst, 100.0 * smp[op] / max(n_smp, 1)) for op, n in mix.most_common(14)) + "\n")
    print("stall reasons: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(n_smp, 1)) for k, v in stalls.most_common(7)) + "\n")
    # ---- CUDA view: per source line
    lines = run(rep, kern, launch, "cuda,sass")
    per = collections.OrderedDict()
    path = ""
    for row in csv.reader(lines):
        if len(row) == 2 and row[0] == "File Path":
            path = row[1].split("/")[-1]
            continue
        if len(row) < 10 or not row[0].isdigit():
            continue                                  # SASS rows have an empty line number
        key = (path, int(row[0]))
        if not (row[7].isdigit() and row[6].isdigit()):
            continue
        inst, s = int(row[7]), int(row[6])
        if key in per:
            per[key][1] += inst; per[key][2] += s
        else:
            per[key] = [row[1].strip(), inst, s]
    tot = sum(v[1] for v in per.values()) or 1
    print("| file:line | source | instructions | share | stall samples |\n|---|---|---:|---:|---:|")
    for (p, ln), (src, inst, s) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
        print("| %s:%d | `%s` | %d | %.1f%% | %d |" % (p, ln, src[:70].replace("|", "\\|"), inst, 100.0 * inst / tot, s))
    print()


if __name__ == "__main__":
    main()
