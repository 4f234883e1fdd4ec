#!/bin/bash
# round 2, GPU call 9: pipelined attention backward (v2) — parity, A/B timing against v1, whole suite, ncu
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_bwd_gpu.py -m gpu -q -s > gpurun_out/r2c9_pytest_attn.log 2>&1; echo "attn bwd pytest rc=$?"; grep -E "rel err|passed|failed" gpurun_out/r2c9_pytest_attn.log | tail -4
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2c9_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/r2c9_pytest.log | tail -3
timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c9_train_b32.json 2> gpurun_out/r2c9_train_b32.err; echo "train32 rc=$?"
HUPR_ATTN_BWD_V1=1 timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c9_train_b32_v1.json 2> gpurun_out/r2c9_train_b32_v1.err; echo "train32 v1 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c9_train_b32_bf16.json 2> gpurun_out/r2c9_train_b32_bf16.err; echo "train32 bf16 rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:attention_bwd --launch-skip 1 -c 1 -o gpurun_out/r02_attnbwd_v2 python tools_dev/prof_kernels.py attnbwd_fused 8 > gpurun_out/r2c9_ncu.log 2>&1; echo "ncu attnbwd v2 rc=$?"
python - <<'PY'
import json
for f in ("r2c9_train_b32","r2c9_train_b32_v1","r2c9_train_b32_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), d["breakdown"].get("attention_bwd"))
    except Exception as e:
        print(f, "ERR", e)
PY
