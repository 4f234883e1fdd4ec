#!/bin/bash
# round 2, GPU call 20: first run of the CTA-pair (cta_group::2) halo convolution — short timeouts, a hang must not eat the box
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_conv_quant_gpu.py -x -q -s -k "two_unit_conv" > gpurun_out/r2c20_pytest_pairs.log 2>&1; echo "pairs pytest rc=$?"; tail -8 gpurun_out/r2c20_pytest_pairs.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
timeout 120 python tools_dev/time_halo_grid.py > gpurun_out/r2c20_halo_grid.txt 2>&1; echo "grid sweep rc=$?"; cat gpurun_out/r2c20_halo_grid.txt
HUPR_HALO_SINGLE=1 timeout 120 python tools_dev/time_halo_grid.py > gpurun_out/r2c20_halo_grid_single.txt 2>&1; echo "grid sweep single rc=$?"; cat gpurun_out/r2c20_halo_grid_single.txt
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_model_gpu.py tests/test_pipeline_gpu.py -x -q > gpurun_out/r2c20_pytest_model.log 2>&1; echo "model pytest rc=$?"; tail -5 gpurun_out/r2c20_pytest_model.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c20_e2e.json 2> gpurun_out/r2c20_e2e.err; echo "e2e rc=$?"
python - <<'PY'
import json
for f in ("r2c20_e2e",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("kernel_ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
