#!/bin/bash
# The N-slab evidence run (N GPUs of one box, under gpurun --gpus N): the slab parity tests that need
# exactly N ranks, the bench line, the in-kernel timeline of a PCG iteration on every rank.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/profile_slabs.sh 8 <tag> [notests]'   -> gpurun_out/<tag>_*
n=${1:-8}
tag=${2:-slabs}
out=gpurun_out
mkdir -p $out
launch="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
if [ "$3" != "notests" ]; then
  ( time timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "test_slabs_match_single_gpu and ${n}-" ) > $out/${tag}_gpu_tests_${n}slabs.log 2>&1
  tail -n 6 $out/${tag}_gpu_tests_${n}slabs.log
fi
timeout 300 $launch --master-port 29531 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu > $out/${tag}_bench_16384_rb_n$n.json 2> $out/${tag}_bench_n$n.err
tail -c 300 $out/${tag}_bench_16384_rb_n$n.json; echo
timeout 200 $launch --master-port 29532 tools/iter_trace.py 16384 16384 2>&1 | grep -v "Warning\|^\*\*\*\|OMP_NUM\|^$" > $out/${tag}_trace_n$n.log
cut -c1-260 $out/${tag}_trace_n$n.log
