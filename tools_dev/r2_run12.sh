#!/bin/bash
# round 2, GPU call 12: attention backward v2 with deferred dQ wait / batched TMEM stores, channel_sums with all loads in flight
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_bwd_gpu.py tests/test_training_blocks_gpu.py -m gpu -q -s > gpurun_out/r2c12_pytest_a.log 2>&1; echo "pytest attn/blocks rc=$?"; grep -E "rel err|passed|failed" gpurun_out/r2c12_pytest_a.log | tail -4
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c12_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/r2c12_pytest.log | tail -3
timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c12_train_b32.json 2> gpurun_out/r2c12_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c12_train_b32_bf16.json 2> gpurun_out/r2c12_train_b32_bf16.err; echo "train32 bf16 rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:attention_bwd --launch-skip 1 -c 1 -o gpurun_out/r02_attnbwd_v3 python tools_dev/prof_kernels.py attnbwd_fused 8 > gpurun_out/r2c12_ncu.log 2>&1; echo "ncu attnbwd rc=$?"
python - <<'PY'
import json
for f in ("r2c12_train_b32","r2c12_train_b32_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        bd=d["breakdown"]
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), bd.get("attention_bwd"), bd.get("channel_sums"))
    except Exception as e:
        print(f, "ERR", e)
PY
