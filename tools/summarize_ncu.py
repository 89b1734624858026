#!/usr/bin/env python
"""Turn ncu outputs under gpurun_out/ into the small text summaries committed under profiles/.

  summarize_ncu.py launches <launches.csv> <out.md>      per-kernel time shares of one step
  summarize_ncu.py full <report.ncu-rep> <out.md>        key metrics of an `ncu --set full` capture
"""
import collections
import csv
import re
import subprocess
import sys


def short(name):
    return re.sub(r"\(.*", "", name).split("::")[-1]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(short(r[ki]), [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# ncu launch list summary (`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("source: %s — %d launches, %.1f ms of kernel time (cold-cache, serialised: compare SHARES)\n\n" % (path, sum(a[0] for a in agg.values()), tot / 1e3))
        f.write("| kernel | launches | total us | avg us | share | grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for k, (n, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.2f | %.1f%% | %s | %s |\n" % (k, n, t, t / n, 100 * t / tot, g, b))
    print("wrote", out)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full summary\n\nsource: %s\n" % path)
        for r in rows[2:]:
            f.write("\n## %s\n\n| metric | value | unit |\n|---|---:|---|\n" % short(r[hdr.index("Kernel Name")]))
            vals = {}
            for m in WANT:
                if m in hdr:
                    i = hdr.index(m)
                    vals[m] = (r[i], units[i])
                    f.write("| %s | %s | %s |\n" % (m, r[i], units[i]))
            try:
                def tobytes(m):
                    v, u = vals[m]
                    v = float(v.replace(",", ""))
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                def tosec(m):
                    v, u = vals[m]
                    v = float(v.replace(",", ""))
                    return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u, 1e-9)
                tr = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
                f.write("| **DRAM traffic per launch** | %.4f | GB |\n| **DRAM GB/s under ncu** | %.0f | GB/s |\n"
                        % (tr / 1e9, tr / tosec("gpu__time_duration.sum") / 1e9))
            except Exception as e:  # pragma: no cover
                f.write("| traffic | n/a (%s) | |\n" % e)
    print("wrote", out)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
