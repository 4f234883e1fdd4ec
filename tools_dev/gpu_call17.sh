#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training_step_gpu.py tests/test_training_blocks_gpu.py tests/test_runner_gpu.py -m gpu -x -q > gpurun_out/c17_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/c17_tests.log; grep -E "^E" gpurun_out/c17_tests.log | head -5
timeout 300 python bench.py --workload train --batch 16 --steps 10 --no-cpu-baseline > gpurun_out/c17_train.json 2>/dev/null; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c17_train.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["value"],1), d["launches_per_step"])
PY
