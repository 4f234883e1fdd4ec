#!/bin/bash
# round 2, GPU call 3: tests after the wgrad producer rework / hi|lo N-concat / carry-counter producers; bench lines; batch-1 launch list
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s --durations=6 > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/r2c3_pytest.log | tail -4
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2c3_e2e.json 2> gpurun_out/r2c3_e2e.err; echo "e2e rc=$?"
timeout 300 python bench.py --workload forward-b1 --steps 200 --no-cpu-baseline > gpurun_out/r2c3_b1.json 2> gpurun_out/r2c3_b1.err; echo "b1 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c3_train_b32.json 2> gpurun_out/r2c3_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c3_train_b32_bf16.json 2> gpurun_out/r2c3_train_b32_bf16.err; echo "train32 bf16 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_b1_launches.csv python bench.py --workload forward-b1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c3_ncu_b1.log 2>&1; echo "b1 list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_train_step_launches_v2.csv python tools_dev/train_one_step.py 32 3 > gpurun_out/r2c3_ncu_train.log 2>&1; echo "train list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_train_step_bf16_launches_v2.csv python tools_dev/train_one_step.py 32 1 > gpurun_out/r2c3_ncu_train_bf16.log 2>&1; echo "train bf16 list rc=$?"
python - <<'PY'
import json
for f in ("r2c3_e2e","r2c3_b1","r2c3_train_b32","r2c3_train_b32_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
