#!/bin/bash
# The N=1 evidence run (one B200, under gpurun): bench line, reference arm, ncu launch list of the
# same bench command, ncu --set full of two consecutive PCG iterations, in-kernel timeline.
#   gpurun --timeout 1500 -- 'bash tools/profile_n1.sh <tag>'     -> gpurun_out/<tag>_*
# Numbers printed by runs under ncu are never bench values (B200_PROFILING.md).
tag=${1:-prof}
out=gpurun_out
mkdir -p $out
python bench.py --gpus 1 --steps 10 --warmup 3 > $out/${tag}_bench_16384_rb.json 2> $out/${tag}_bench.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench.err
# launch list: one warm-up and two timed sub-steps, no CPU arm, no tolerance study
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
    python bench.py --gpus 1 --steps 2 --warmup 1 --no-cpu --no-tol-study --no-kernel-timers > $out/${tag}_ncu_bench.log 2>&1
# full capture: PCG iterations 10 and 11 of a solve (an even and an odd tail launch)
ncu --set full --clock-control none --import-source on -k regex:'k_fused_tail|k_fused_search_apply' -s 18 -c 4 \
    -f -o $out/${tag}_pcg python tools/perf_probe.py 16384 rb 1 > $out/${tag}_ncu_full.log 2>&1
python tools/iter_trace.py 16384 16384 > $out/${tag}_trace_full_n1.log 2>&1
python tools/iter_trace.py 16384 2150 > $out/${tag}_trace_thin_n1.log 2>&1
tail -c 600 $out/${tag}_bench_16384_rb.json; echo; ls -la $out | grep ${tag}_
