#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --workload train --batch 16 --steps 10 --no-cpu-baseline > gpurun_out/c14_train_nfast.json 2>/dev/null; echo "rc=$?"
HUPR_M_FASTEST=1 timeout 300 python bench.py --workload train --batch 16 --steps 10 --no-cpu-baseline > gpurun_out/c14_train_mfast.json 2>/dev/null; echo "rc=$?"
python - <<'PY'
import json
for f in ("c14_train_nfast","c14_train_mfast"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["ms_per_step"],3), {k:v["ms"] for k,v in list(d["breakdown"]["conv_gemm_by_shape"].items())[:4]}, d["breakdown"]["matmul_tn"])
PY
