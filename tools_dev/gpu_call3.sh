#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training_step_gpu.py tests/test_training_blocks_gpu.py tests/test_runner_gpu.py tests/test_wgrad_gpu.py -m gpu -x -q > gpurun_out/c3_train_tests.log 2>&1
echo "training tests rc=$?"
tail -5 gpurun_out/c3_train_tests.log
timeout 600 python bench.py --workload train --batch 16 --steps 10 --no-cpu-baseline > gpurun_out/c3_bench_train_b16.json 2> gpurun_out/c3_bench_train_b16.err
echo "bench rc=$?"
tail -3 gpurun_out/c3_bench_train_b16.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c3_bench_train_b16.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k,v in d["breakdown"].items():
    if k!="conv_gemm_by_shape": print(" ",k,v)
for k,v in d["breakdown"]["conv_gemm_by_shape"].items(): print("    ",k,v)
PY
