#!/bin/bash
# round-1 re-entry check: parity of HEAD (attention P-in-TMEM), full gpu suite, e2e bench for both attention variants
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -x -q > gpurun_out/c1_model_tmem.log 2>&1
echo "model tests (P in TMEM) rc=$?" | tee -a gpurun_out/c1_summary.txt
timeout 200 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/c1_bench_e2e_tmem.json 2> gpurun_out/c1_bench_e2e_tmem.err
echo "bench tmem rc=$?" | tee -a gpurun_out/c1_summary.txt
HUPR_NVCC_EXTRA=-DHUPR_ATTN_P_IN_TMEM=0 python -m hupr_b200.build --force > gpurun_out/c1_rebuild.log 2>&1
echo "rebuild smem rc=$?" | tee -a gpurun_out/c1_summary.txt
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -x -q > gpurun_out/c1_model_smem.log 2>&1
echo "model tests (P in smem) rc=$?" | tee -a gpurun_out/c1_summary.txt
timeout 200 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/c1_bench_e2e_smem.json 2> gpurun_out/c1_bench_e2e_smem.err
echo "bench smem rc=$?" | tee -a gpurun_out/c1_summary.txt
# back to the default build for the rest of the suite
python -m hupr_b200.build --force > gpurun_out/c1_rebuild2.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/c1_pytest_all.log 2>&1
echo "full gpu suite rc=$?" | tee -a gpurun_out/c1_summary.txt
tail -3 gpurun_out/c1_pytest_all.log
