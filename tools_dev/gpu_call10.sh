#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2; do
timeout 200 python bench.py --workload forward-b1 --steps 100 --no-cpu-baseline > gpurun_out/c10_b1_coop_$i.json 2>/dev/null
HUPR_NO_COOP=1 timeout 200 python bench.py --workload forward-b1 --steps 100 --no-cpu-baseline > gpurun_out/c10_b1_nocoop_$i.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c10_b1_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],4), round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], {k:v["ms"] for k,v in d["breakdown"]["conv_gemm_by_shape"].items()})
PY
