#!/bin/bash
# round 2, GPU call 19: two-unit convolution tests (all), per-SM vs chip-wide bound of the halo kernels, default-path bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_quant_gpu.py -q -s > gpurun_out/r2c19_pytest_quant.log 2>&1; echo "quant pytest rc=$?"; tail -12 gpurun_out/r2c19_pytest_quant.log
timeout 300 python tools_dev/time_halo_grid.py > gpurun_out/r2c19_halo_grid.txt 2>&1; echo "grid sweep rc=$?"; cat gpurun_out/r2c19_halo_grid.txt
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_model_gpu.py tests/test_conv_gemm_gpu.py tests/test_training_step_gpu.py -q -s > gpurun_out/r2c19_pytest_model.log 2>&1; echo "model pytest rc=$?"; grep -i "two-unit\|passed\|failed\|error" gpurun_out/r2c19_pytest_model.log | tail -8
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c19_e2e.json 2> gpurun_out/r2c19_e2e.err; echo "e2e rc=$?"
timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c19_train_b32.json 2> gpurun_out/r2c19_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c19_train_b32_bf16.json 2> gpurun_out/r2c19_train_b32_bf16.err; echo "train32 bf16 rc=$?"
HUPR_QUANT=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c19_e2e_quant.json 2> gpurun_out/r2c19_e2e_quant.err; echo "e2e quant rc=$?"
timeout 300 python bench.py --workload forward-b1 --steps 200 --no-cpu-baseline > gpurun_out/r2c19_b1.json 2> gpurun_out/r2c19_b1.err; echo "b1 rc=$?"
python - <<'PY'
import json
for f in ("r2c19_e2e","r2c19_e2e_quant","r2c19_b1","r2c19_train_b32","r2c19_train_b32_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("kernel_ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
