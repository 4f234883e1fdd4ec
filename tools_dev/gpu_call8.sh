#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python bench.py --workload train --batch 32 --steps 5 --no-cpu-baseline > gpurun_out/c8_train_b32.json 2> gpurun_out/c8_train_b32.err; echo "train b32 rc=$?"; tail -2 gpurun_out/c8_train_b32.err; cut -c1-300 gpurun_out/c8_train_b32.json
python - <<'PY'
import torch
print("peak mem note: run separately")
PY
timeout 300 python bench.py --workload forward-b1 --steps 50 > gpurun_out/c8_b1.json 2> gpurun_out/c8_b1.err; echo "b1 rc=$?"; cut -c1-300 gpurun_out/c8_b1.json
timeout 300 python bench.py --workload cascade --steps 20 > gpurun_out/c8_cascade.json 2> gpurun_out/c8_cascade.err; echo "cascade rc=$?"; cut -c1-300 gpurun_out/c8_cascade.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/c8_ref.json 2> gpurun_out/c8_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/c8_ref.json
