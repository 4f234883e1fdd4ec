#!/bin/bash
# round 2, GPU call 21: CTA pairs in all three arithmetic modes — whole GPU suite, benches, grid sweep
set -u
mkdir -p gpurun_out
timeout 120 python tools_dev/time_halo_grid.py > gpurun_out/r2c21_halo_grid.txt 2>&1; echo "grid sweep rc=$?"; cat gpurun_out/r2c21_halo_grid.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2c21_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2c21_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c21_e2e.json 2> gpurun_out/r2c21_e2e.err; echo "e2e rc=$?"
HUPR_QUANT=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c21_e2e_quant.json 2> gpurun_out/r2c21_e2e_quant.err; echo "e2e quant rc=$?"
HUPR_QUANT=1 HUPR_QUANT_FUSE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c21_e2e_quant_nofuse.json 2> gpurun_out/r2c21_e2e_quant_nofuse.err; echo "e2e quant nofuse rc=$?"
timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c21_train_b32.json 2> gpurun_out/r2c21_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c21_train_b32_bf16.json 2> gpurun_out/r2c21_train_b32_bf16.err; echo "train32 bf16 rc=$?"
python - <<'PY'
import json
for f in ("r2c21_e2e","r2c21_e2e_quant","r2c21_e2e_quant_nofuse","r2c21_train_b32","r2c21_train_b32_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("kernel_ms_per_step"), r.get("executed_tensor_tflops"))
        for k,v in list(d["breakdown"]["conv_gemm_by_shape"].items())[:4]: print("      ",k,v)
    except Exception as e:
        print(f, "ERR", e)
PY
