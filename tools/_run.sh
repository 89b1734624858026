( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/t31_all_1gpu.log 2>&1
tail -n 8 gpurun_out/t31_all_1gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-tol-study > gpurun_out/b31_n1.json 2> gpurun_out/b31_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b31_n1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"], d["roofline"])
PY
