#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest_all.log 2>&1
echo "full gpu suite rc=$?"; tail -3 gpurun_out/c6_pytest_all.log
timeout 300 python bench.py --steps 20 > gpurun_out/c6_bench_e2e.json 2> gpurun_out/c6_bench_e2e.err
echo "bench e2e rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c6_bench_e2e.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["clocks"])
for k,v in d["breakdown"].items():
    if k!="conv_gemm_by_shape": print(" ",k,v)
for k,v in d["breakdown"]["conv_gemm_by_shape"].items(): print("    ",k,v)
PY
