#!/bin/bash
# round 2, GPU call 30: saturation counter of the two-unit operand planes, cluster fallback query, whole suite + smoke + default bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c30_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c30_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c30_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2c30_smoke.log
timeout 400 python bench.py > gpurun_out/r2c30_e2e.json 2> gpurun_out/r2c30_e2e.err; echo "e2e rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c30_e2e.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],3), d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["cpu_baseline"]["value"])
PY
