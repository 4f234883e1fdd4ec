#!/bin/bash
# round 2, GPU call 29: compute-sanitizer over the CTA-pair / two-unit halo kernels and the reworked element-wise training kernels
set -u
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_conv_quant_gpu.py tests/test_conv_gemm_gpu.py -x -q > gpurun_out/r02_sanitizer_memcheck_pairs.log 2>&1; echo "memcheck pairs rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck_pairs.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_conv_quant_gpu.py -x -q -k "two_unit_conv or fall_back" > gpurun_out/r02_sanitizer_racecheck_pairs.log 2>&1; echo "racecheck pairs rc=$?"; tail -4 gpurun_out/r02_sanitizer_racecheck_pairs.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_training_blocks_gpu.py -x -q > gpurun_out/r02_sanitizer_memcheck_train_blocks.log 2>&1; echo "memcheck train blocks rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck_train_blocks.log
