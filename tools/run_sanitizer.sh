#!/bin/bash
# compute-sanitizer passes over the GPU parity tests on small grids (the new build's
# replacement for the reference's SHERLOCK FP traps).  Run on a GPU box from the repo root:
#     bash tools/run_sanitizer.sh [out_dir]
# memcheck: out-of-bounds / misaligned accesses (the TMA ring, halo columns, guard rows);
# racecheck: shared-memory hazards (the mbarrier ring of pcg_pipe.cuh, the rolling neighbour buffer
#            of pcg_tail.cuh); initcheck: reads of uninitialised device memory.
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SAN=/usr/local/cuda/bin/compute-sanitizer
SMALL='(fused_tail or pressure_solve_pieces or marker_and_grid) and (block-100-40 or block-64-48 or waterfall-100-40)'
FRAMES='test_frames_red_black and block or test_tile_list and waterfall'
for tool in memcheck racecheck initcheck; do
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check no"
  [ "$tool" = racecheck ] && extra="--racecheck-report all --print-limit 4000"
  timeout 1500 $SAN --tool $tool --error-exitcode 86 --print-limit 20 $extra \
    python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "$SMALL" > "$OUT/sanitizer_${tool}_stages.log" 2>&1
  echo "$tool stages rc=$?" >> "$OUT/sanitizer_summary.log"
  tail -4 "$OUT/sanitizer_${tool}_stages.log" >> "$OUT/sanitizer_summary.log"
done
# racecheck does not model mbarrier arrive / wait as the synchronisation between the consumers'
# reads of a ring stage and the bulk copy that refills it (generic-proxy read vs async-proxy write:
# the standard TMA pipeline hand-over), so it reports every such pair as a "potential WAR hazard".
# Pair up the read and write sites of everything it printed: anything that is NOT (a read of a ring
# stage, pipe::bulk_g2s) would be a real finding.
python - "$OUT/sanitizer_racecheck_stages.log" >> "$OUT/sanitizer_summary.log" <<'PY'
import collections, re, sys
pairs, kind = collections.Counter(), None
rd = None
for line in open(sys.argv[1], errors="replace"):
    m = re.search(r"(Error|Warning): (.*?) detected", line)
    if m: kind = m.group(2)
    m = re.search(r"(Read|Write) Thread .* at (.*?)\+0x[0-9a-f]+ in (\S+)", line)
    if m:
        site = "%s (%s)" % (re.sub(r"\(.*", "", m.group(2)).split("::")[-1], m.group(3))
        if m.group(1) == "Read": rd = site
        else: pairs[(kind, rd, site)] += 1
print("racecheck hazards printed, by (kind, read site, write site):")
for (k, r, w), n in pairs.most_common():
    print("  %6d  %s: read %s / write %s" % (n, k, r, w))
other = [x for x in pairs if "bulk_g2s" not in x[2]]
print("hazards that do not involve the bulk copy refilling a ring stage:", len(other))
PY
timeout 1500 $SAN --tool memcheck --leak-check no --error-exitcode 86 --print-limit 20 \
  python -m pytest tests/test_gpu_frames.py -m gpu -x -q -k "$FRAMES" > "$OUT/sanitizer_memcheck_frames.log" 2>&1
echo "memcheck frames rc=$?" >> "$OUT/sanitizer_summary.log"
tail -4 "$OUT/sanitizer_memcheck_frames.log" >> "$OUT/sanitizer_summary.log"
cat "$OUT/sanitizer_summary.log"
