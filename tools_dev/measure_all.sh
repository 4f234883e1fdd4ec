#!/bin/bash
# round-end measurement batch: tests, smoke, every bench line, e2e launch list (ncu numbers are never bench values)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/f_smoke.log
timeout 400 python bench.py > gpurun_out/f_e2e.json 2> gpurun_out/f_e2e.err; echo "e2e rc=$?"
timeout 400 python bench.py --workload train --steps 10 > gpurun_out/f_train_b32.json 2> gpurun_out/f_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --batch 16 --steps 10 --no-cpu-baseline > gpurun_out/f_train_b16.json 2> gpurun_out/f_train_b16.err; echo "train16 rc=$?"
timeout 300 python bench.py --workload forward-b1 --steps 100 > gpurun_out/f_b1.json 2> gpurun_out/f_b1.err; echo "b1 rc=$?"
timeout 300 python bench.py --workload cascade > gpurun_out/f_cascade.json 2> gpurun_out/f_cascade.err; echo "cascade rc=$?"
timeout 300 python bench.py --workload cascade-sweep > gpurun_out/f_sweep.json 2> gpurun_out/f_sweep.err; echo "sweep rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; echo "ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_e2e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_e2e.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json
for f in ("f_e2e","f_train_b32","f_train_b16","f_b1","f_cascade","f_sweep","f_ref"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
