#!/bin/bash
# round 2, GPU call 31: last validation of the final build + refreshed training bench lines and launch lists
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/i2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/i2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/i2_smoke.log
timeout 400 python bench.py > gpurun_out/i2_e2e.json 2> gpurun_out/i2_e2e.err; echo "e2e rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/i2_ref.json 2> gpurun_out/i2_ref.err; echo "ref rc=$?"
timeout 400 python bench.py --workload train --steps 10 > gpurun_out/i2_train_b32.json 2> gpurun_out/i2_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/i2_train_b32_bf16.json 2> gpurun_out/i2_train_b32_bf16.err; echo "train32 bf16 rc=$?"
timeout 400 python bench.py --workload train --batch 16 --steps 10 --no-cpu-baseline > gpurun_out/i2_train_b16.json 2> gpurun_out/i2_train_b16.err; echo "train16 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_train_step_launches_v5.csv python tools_dev/train_one_step.py 32 3 > gpurun_out/i2_ncu_train.log 2>&1; echo "train list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_train_step_bf16_launches_v5.csv python tools_dev/train_one_step.py 32 1 > gpurun_out/i2_ncu_train_bf16.log 2>&1; echo "train bf16 list rc=$?"
python - <<'PY'
import json
for f in ("i2_e2e","i2_ref","i2_train_b32","i2_train_b32_bf16","i2_train_b16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), r.get("frac"), r.get("traffic"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY
