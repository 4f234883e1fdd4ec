#!/bin/bash
# round 2, GPU call 33 (gpurun --gpus 4): the default multi-rank bench line (stream + train_probe) on 4 ranks, final build
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29551 bench.py --gpus 4 > gpurun_out/k2_e2e_4gpu.json 2> gpurun_out/k2_e2e_4gpu.err; echo "e2e 4gpu rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/k2_e2e_4gpu.json").read().strip().splitlines()[-1])
print(round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), d.get("train_probe"), d.get("clocks"))
PY
