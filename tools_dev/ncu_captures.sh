#!/bin/bash
# ncu captures (never a bench value): direct wgrad kernel, attention with P in TMEM, persistent halo conv; launch list of one training step
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:wgrad_kernel -c 3 -o gpurun_out/r01_wgrad python tools_dev/prof_kernels.py wgrad 16 > gpurun_out/c4_ncu_wgrad.log 2>&1; echo "wgrad ncu rc=$?"
timeout 400 $NCU -k regex:attention_kernel --launch-skip 1 -c 1 -o gpurun_out/r01_attn_v4 python tools_dev/prof_kernels.py attn 32 > gpurun_out/c4_ncu_attn.log 2>&1; echo "attn ncu rc=$?"
timeout 400 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r01_halo128_persist python tools_dev/prof_kernels.py conv128 32 > gpurun_out/c4_ncu_halo.log 2>&1; echo "halo ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r01_train_launches.csv python bench.py --workload train --batch 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c4_ncu_train.log 2>&1; echo "train list rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r01_train_launches.csv
