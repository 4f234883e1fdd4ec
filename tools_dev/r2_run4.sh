#!/bin/bash
# round 2, GPU call 4 (gpurun --gpus 2): data-parallel training on hardware — 2-rank NCCL gradient parity test, the default multi-rank
# bench line with its train_probe record, and the training workload itself on two ranks (3-product and bf16).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2c4_smi.txt 2>&1
timeout 900 python -m pytest tests/test_dp_training_gpu.py -m gpu -q -s > gpurun_out/r2c4_pytest.log 2>&1; echo "dp pytest rc=$?"; tail -5 gpurun_out/r2c4_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c4_e2e_2gpu.json 2> gpurun_out/r2c4_e2e_2gpu.err; echo "e2e 2gpu rc=$?"
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c4_train_2gpu.json 2> gpurun_out/r2c4_train_2gpu.err; echo "train 2gpu rc=$?"
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c4_train_2gpu_bf16.json 2> gpurun_out/r2c4_train_2gpu_bf16.err; echo "train 2gpu bf16 rc=$?"
python - <<'PY'
import json
for f in ("r2c4_e2e_2gpu","r2c4_train_2gpu","r2c4_train_2gpu_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), d.get("train_probe"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2c4_*.err
