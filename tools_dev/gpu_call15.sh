#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_training_blocks_gpu.py tests/test_model_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/c15_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/c15_tests.log; grep -E "^E" gpurun_out/c15_tests.log | head -5
timeout 300 python bench.py --workload train --batch 16 --steps 10 --no-cpu-baseline > gpurun_out/c15_train.json 2>/dev/null; echo "rc=$?"
timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/c15_e2e.json 2>/dev/null; echo "rc=$?"
timeout 300 python bench.py --workload forward-b1 --steps 100 --no-cpu-baseline > gpurun_out/c15_b1.json 2>/dev/null; echo "rc=$?"
python - <<'PY'
import json
for f in ("c15_train","c15_e2e","c15_b1"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["value"],1), {k:v["ms"] for k,v in list(d["breakdown"]["conv_gemm_by_shape"].items())[:5]})
PY
