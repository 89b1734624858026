timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/iter_trace.py 16384 4300 > gpurun_out/tr30_thin_n2_final.log 2>&1
grep -A2 "^rank" gpurun_out/tr30_thin_n2_final.log | cut -c1-330
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_stages.py tests/test_gpu_converged.py -m gpu -x -q ) > gpurun_out/t30_multi_stages_converged_2gpu.log 2>&1
tail -n 8 gpurun_out/t30_multi_stages_converged_2gpu.log
