#!/bin/bash
# round 2, GPU call 27: CTA pairs for the cout = 64 single-product kernel; saturation / fallback test of the two-unit convolution
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_conv_quant_gpu.py tests/test_training_step_gpu.py tests/test_training_blocks_gpu.py -q -s > gpurun_out/r2c27_pytest.log 2>&1; echo "pytest rc=$?"; grep -i "up to\|passed\|failed" gpurun_out/r2c27_pytest.log | tail -5
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c27_train_b32_bf16.json 2> gpurun_out/r2c27_train_b32_bf16.err; echo "train32 bf16 rc=$?"
python - <<'PY'
import json
for f in ("r2c27_train_b32_bf16",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("kernel_ms_per_step"))
        for k,v in list(d["breakdown"]["conv_gemm_by_shape"].items())[:5]: print("      ",k,v)
    except Exception as e:
        print(f, "ERR", e)
PY
