#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest_all.log 2>&1
echo "full gpu suite rc=$?"; tail -4 gpurun_out/c11_pytest_all.log; grep -E "^E" gpurun_out/c11_pytest_all.log | head
