#!/bin/bash
# 8-GPU lines: e2e stream, training step (batch 32 per GPU = BASELINE configs[3]: 256 over 8), cascade sweep
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 200 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/c16_e2e_8gpu.json 2> gpurun_out/c16_e2e_8gpu.err; echo "e2e rc=$?"; cut -c1-260 gpurun_out/c16_e2e_8gpu.json
timeout 200 $TR bench.py --gpus 8 --workload train --steps 5 --warmup 3 > gpurun_out/c16_train_8gpu.json 2> gpurun_out/c16_train_8gpu.err; echo "train rc=$?"; cut -c1-260 gpurun_out/c16_train_8gpu.json
timeout 120 $TR bench.py --gpus 8 --workload cascade-sweep > gpurun_out/c16_sweep_8gpu.json 2> gpurun_out/c16_sweep_8gpu.err; echo "sweep rc=$?"; cut -c1-200 gpurun_out/c16_sweep_8gpu.json
