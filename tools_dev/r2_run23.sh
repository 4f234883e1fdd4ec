#!/bin/bash
# round 2, GPU call 23: CTA pairs for the cout = 64 halo kernel
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gemm_gpu.py tests/test_conv_quant_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/r2c23_pytest_a.log 2>&1; echo "pytest a rc=$?"; tail -3 gpurun_out/r2c23_pytest_a.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c23_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c23_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c23_e2e.json 2> gpurun_out/r2c23_e2e.err; echo "e2e rc=$?"
timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c23_train_b32.json 2> gpurun_out/r2c23_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c23_train_b32_bf16.json 2> gpurun_out/r2c23_train_b32_bf16.err; echo "train32 bf16 rc=$?"
python - <<'PY'
import json
for f in ("r2c23_e2e","r2c23_train_b32","r2c23_train_b32_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("kernel_ms_per_step"))
        for k,v in list(d["breakdown"]["conv_gemm_by_shape"].items())[:5]: print("      ",k,v)
    except Exception as e:
        print(f, "ERR", e)
PY
