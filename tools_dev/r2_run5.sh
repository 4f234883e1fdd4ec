#!/bin/bash
# round 2, GPU call 5: batched pack/unpack launches — tests, training bench lines, kernel census
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2c5_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/r2c5_pytest.log | tail -4
timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c5_train_b32.json 2> gpurun_out/r2c5_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c5_train_b32_bf16.json 2> gpurun_out/r2c5_train_b32_bf16.err; echo "train32 bf16 rc=$?"
timeout 300 python bench.py --steps 30 > gpurun_out/r2c5_e2e.json 2> gpurun_out/r2c5_e2e.err; echo "e2e rc=$?"
python - <<'PY'
import json
for f in ("r2c5_e2e","r2c5_train_b32","r2c5_train_b32_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("kernels_per_replay"), d.get("launches_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2c5_*.err
