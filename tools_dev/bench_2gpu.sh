#!/bin/bash
# 2-GPU lines: e2e stream, training step (NCCL gradient all-reduce), cascade sweep
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c7_e2e_2gpu.json 2> gpurun_out/c7_e2e_2gpu.err; echo "e2e rc=$?"; cut -c1-400 gpurun_out/c7_e2e_2gpu.json
timeout 300 $TR bench.py --gpus 2 --workload train --batch 16 --steps 5 --warmup 3 > gpurun_out/c7_train_2gpu.json 2> gpurun_out/c7_train_2gpu.err; echo "train rc=$?"; cut -c1-400 gpurun_out/c7_train_2gpu.json
timeout 300 $TR bench.py --gpus 2 --workload cascade-sweep > gpurun_out/c7_sweep_2gpu.json 2> gpurun_out/c7_sweep_2gpu.err; echo "sweep rc=$?"; cut -c1-300 gpurun_out/c7_sweep_2gpu.json
tail -3 gpurun_out/c7_*.err
