#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:conv_gemm_kernel --launch-skip 2 -c 2 -o gpurun_out/r01_attnbwd_gemm python tools_dev/prof_kernels.py attnbwd 16 > gpurun_out/c13_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/c13_ncu.log
