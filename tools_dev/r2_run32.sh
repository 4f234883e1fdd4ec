#!/bin/bash
# round 2, GPU call 32 (gpurun --gpus 2): final build on 2 ranks — stream + train_probe, training workload (3 products, bf16), DP parity test
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29541 bench.py --gpus 2 > gpurun_out/j2_e2e_2gpu.json 2> gpurun_out/j2_e2e_2gpu.err; echo "e2e 2gpu rc=$?"
timeout 500 $TR --master-port 29542 bench.py --gpus 2 --workload train --steps 10 --no-cpu-baseline > gpurun_out/j2_train_2gpu.json 2> gpurun_out/j2_train_2gpu.err; echo "train 2gpu rc=$?"
timeout 500 $TR --master-port 29543 bench.py --gpus 2 --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/j2_train_2gpu_bf16.json 2> gpurun_out/j2_train_2gpu_bf16.err; echo "train 2gpu bf16 rc=$?"
timeout 600 python -m pytest tests/test_dp_training_gpu.py -m gpu -q > gpurun_out/j2_pytest_dp.log 2>&1; echo "dp pytest rc=$?"; tail -2 gpurun_out/j2_pytest_dp.log
python - <<'PY'
import json
for f in ("j2_e2e_2gpu","j2_train_2gpu","j2_train_2gpu_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), d.get("train_probe"))
    except Exception as e:
        print(f, "ERR", e)
PY
