#!/bin/bash
# round-2 FINAL measurement batch (1 GPU): tests, smoke, every bench line, launch lists (ncu numbers are never bench values)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/h2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/h2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h2_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/h2_smoke.log
timeout 400 python bench.py > gpurun_out/h2_e2e.json 2> gpurun_out/h2_e2e.err; echo "e2e (default flags) rc=$?"
HUPR_QUANT=0 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/h2_e2e_3prod.json 2> gpurun_out/h2_e2e_3prod.err; echo "e2e 3prod rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/h2_ref.json 2> gpurun_out/h2_ref.err; echo "ref rc=$?"
timeout 400 python bench.py --workload train --steps 10 > gpurun_out/h2_train_b32.json 2> gpurun_out/h2_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/h2_train_b32_bf16.json 2> gpurun_out/h2_train_b32_bf16.err; echo "train32 bf16 rc=$?"
timeout 400 python bench.py --workload train --batch 16 --steps 10 --no-cpu-baseline > gpurun_out/h2_train_b16.json 2> gpurun_out/h2_train_b16.err; echo "train16 rc=$?"
timeout 300 python bench.py --workload forward-b1 --steps 200 > gpurun_out/h2_b1.json 2> gpurun_out/h2_b1.err; echo "b1 rc=$?"
timeout 300 python bench.py --workload cascade > gpurun_out/h2_cascade.json 2> gpurun_out/h2_cascade.err; echo "cascade rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_e2e_launches_v3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/h2_ncu_e2e.log 2>&1; echo "e2e list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_train_step_launches_v4.csv python tools_dev/train_one_step.py 32 3 > gpurun_out/h2_ncu_train.log 2>&1; echo "train list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_train_step_bf16_launches_v4.csv python tools_dev/train_one_step.py 32 1 > gpurun_out/h2_ncu_train_bf16.log 2>&1; echo "train bf16 list rc=$?"
python - <<'PY'
import json
for f in ("h2_e2e","h2_e2e_3prod","h2_ref","h2_train_b32","h2_train_b32_bf16","h2_train_b16","h2_b1","h2_cascade"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY
