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
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  timeout 1500 $SAN --tool $tool $extra --error-exitcode 86 --print-limit 20 \
    python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "$SMALL" > "$OUT/sanitizer_${tool}_stages.log" 2>&1
  echo "$tool stages rc=$?" >> "$OUT/sanitizer_summary.log"
  tail -4 "$OUT/sanitizer_${tool}_stages.log" >> "$OUT/sanitizer_summary.log"
done
timeout 1500 $SAN --tool memcheck --leak-check no --error-exitcode 86 --print-limit 20 \
  python -m pytest tests/test_gpu_frames.py -m gpu -x -q -k "$FRAMES" > "$OUT/sanitizer_memcheck_frames.log" 2>&1
echo "memcheck frames rc=$?" >> "$OUT/sanitizer_summary.log"
tail -4 "$OUT/sanitizer_memcheck_frames.log" >> "$OUT/sanitizer_summary.log"
cat "$OUT/sanitizer_summary.log"
