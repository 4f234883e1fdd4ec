#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c9_pytest_all.log 2>&1
echo "full gpu suite rc=$?"; tail -4 gpurun_out/c9_pytest_all.log
timeout 300 python bench.py --workload forward-b1 --steps 50 --no-cpu-baseline > gpurun_out/c9_b1.json 2> gpurun_out/c9_b1.err; echo "b1 rc=$?"
timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/c9_e2e.json 2> gpurun_out/c9_e2e.err; echo "e2e rc=$?"
python - <<'PY'
import json
for f in ("c9_b1","c9_e2e"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"])
    for k,v in d["breakdown"]["conv_gemm_by_shape"].items(): print("    ",k,v)
    for k,v in d["breakdown"].items():
        if k.startswith("gcn") or k=="attention_fwd": print("  ",k,v)
PY
