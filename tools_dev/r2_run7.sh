#!/bin/bash
# round 2, GPU call 7 (gpurun --gpus 8): BASELINE.json configs[3] on the full box — default multi-rank bench line with its train_probe,
# the training workload on 8 ranks (256 samples per step; 3-product and bf16) from pinned host buffers, and the 2-rank NCCL parity test.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2c7_e2e_8gpu.json 2> gpurun_out/r2c7_e2e_8gpu.err; echo "e2e 8gpu rc=$?"
timeout 500 $TR --master-port 29522 bench.py --gpus 8 --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c7_train_8gpu.json 2> gpurun_out/r2c7_train_8gpu.err; echo "train 8gpu rc=$?"
timeout 500 $TR --master-port 29523 bench.py --gpus 8 --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c7_train_8gpu_bf16.json 2> gpurun_out/r2c7_train_8gpu_bf16.err; echo "train 8gpu bf16 rc=$?"
timeout 600 python -m pytest tests/test_dp_training_gpu.py -m gpu -q -s > gpurun_out/r2c7_pytest_dp.log 2>&1; echo "dp pytest rc=$?"; tail -3 gpurun_out/r2c7_pytest_dp.log
python - <<'PY'
import json
for f in ("r2c7_e2e_8gpu","r2c7_train_8gpu","r2c7_train_8gpu_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), d.get("train_probe"))
    except Exception as e:
        print(f, "ERR", e)
PY
