set -x
python -m pytest tests/test_gpu_stages.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t16_stages.log
for th in 32 64 128; do EULER_PCG_TH=$th python tools/slab_probe.py 16384 16384 100 2>&1 | head -6 > gpurun_out/p16_full_th$th.log; done
for th in 32 64 128; do EULER_PCG_TH=$th python tools/slab_probe.py 16384 2150 100 2>&1 | head -6 > gpurun_out/p16_thin_th$th.log; done
cat gpurun_out/t16_stages.log gpurun_out/p16_*.log
