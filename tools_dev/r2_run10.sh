#!/bin/bash
# round 2, GPU call 10: single-thread roles entered through elect.sync (no per-instruction warp-serialisation loops) — tests + every bench line
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2c10_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/r2c10_pytest.log | tail -3
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2c10_e2e.json 2> gpurun_out/r2c10_e2e.err; echo "e2e rc=$?"
timeout 300 python bench.py --workload forward-b1 --steps 200 --no-cpu-baseline > gpurun_out/r2c10_b1.json 2> gpurun_out/r2c10_b1.err; echo "b1 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c10_train_b32.json 2> gpurun_out/r2c10_train_b32.err; echo "train32 rc=$?"
HUPR_ATTN_BWD_V1=1 timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c10_train_b32_v1.json 2> gpurun_out/r2c10_train_b32_v1.err; echo "train32 v1 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c10_train_b32_bf16.json 2> gpurun_out/r2c10_train_b32_bf16.err; echo "train32 bf16 rc=$?"
timeout 300 python bench.py --workload cascade --no-cpu-baseline > gpurun_out/r2c10_cascade.json 2> gpurun_out/r2c10_cascade.err; echo "cascade rc=$?"
python - <<'PY'
import json
for f in ("r2c10_e2e","r2c10_b1","r2c10_train_b32","r2c10_train_b32_v1","r2c10_train_b32_bf16","r2c10_cascade"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("breakdown") or {}).get("attention_bwd"), (d.get("breakdown") or {}).get("attention_fwd"))
    except Exception as e:
        print(f, "ERR", e)
PY
