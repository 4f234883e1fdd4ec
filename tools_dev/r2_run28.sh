#!/bin/bash
# round 2, GPU call 28: training element-wise kernels (pipelined channel_sums, vector loads in mnet_bwd, vector reductions in resample_bwd)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_blocks_gpu.py tests/test_training_step_gpu.py tests/test_model_gpu.py -q > gpurun_out/r2c28_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c28_pytest.log
timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c28_train_b32.json 2> gpurun_out/r2c28_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c28_train_b32_bf16.json 2> gpurun_out/r2c28_train_b32_bf16.err; echo "train32 bf16 rc=$?"
python - <<'PY'
import json
for f in ("r2c28_train_b32","r2c28_train_b32_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"))
        print("   ", {k:v["ms"] for k,v in d["breakdown"].items() if k in ("channel_sums","mnet_bwd","resample_linear_bwd","bn_bwd_apply","affine_act","accumulate","conv_wgrad","attention_bwd")})
    except Exception as e:
        print(f, "ERR", e)
PY
