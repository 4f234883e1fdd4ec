#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ingest_gpu.py tests/test_keypoint_eval.py tests/test_runner_gpu.py tests/test_cascade_gpu.py -m gpu -x -q > gpurun_out/c5_tests.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/c5_tests.log
timeout 300 python bench.py --workload cascade-sweep > gpurun_out/c5_bench_sweep.json 2> gpurun_out/c5_bench_sweep.err
echo "sweep rc=$?"; tail -3 gpurun_out/c5_bench_sweep.err; cat gpurun_out/c5_bench_sweep.json | cut -c1-1500
