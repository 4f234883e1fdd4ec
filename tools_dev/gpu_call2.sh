#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_wgrad_gpu.py -m gpu -x -q > gpurun_out/c2_wgrad_test.log 2>&1
echo "wgrad tests rc=$?"
tail -15 gpurun_out/c2_wgrad_test.log
timeout 300 python tools_dev/time_wgrad.py > gpurun_out/c2_time_wgrad.log 2>&1
echo "time rc=$?"
tail -20 gpurun_out/c2_time_wgrad.log
