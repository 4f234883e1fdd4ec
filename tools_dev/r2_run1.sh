#!/bin/bash
# round 2, GPU call 1: whole gpu test suite, bench lines (e2e, train 3-product / bf16 / unfused-attention A/B, b1), launch lists and
# --set full captures of the HBM-bound kernels.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
# new / reworked areas first, one process per file (a sticky CUDA error must not take the other files down with it)
: > gpurun_out/r2c1_pytest.log
for f in test_pack_gpu test_conv_gemm_gpu test_wgrad_gpu test_training_blocks_gpu test_training_step_gpu test_pipeline_gpu; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -s --durations=8 >> gpurun_out/r2c1_pytest.log 2>&1; echo "pytest $f rc=$?"
done
timeout 900 python -m pytest tests -m gpu -q -s --durations=8 --deselect tests/test_pack_gpu.py --deselect tests/test_conv_gemm_gpu.py --deselect tests/test_wgrad_gpu.py --deselect tests/test_training_blocks_gpu.py --deselect tests/test_training_step_gpu.py --deselect tests/test_pipeline_gpu.py >> gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rest rc=$?"
grep -E "passed|failed|error" gpurun_out/r2c1_pytest.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c1_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2c1_smoke.log
timeout 400 python bench.py > gpurun_out/r2c1_e2e.json 2> gpurun_out/r2c1_e2e.err; echo "e2e rc=$?"
timeout 400 python bench.py --workload train --steps 10 > gpurun_out/r2c1_train_b32.json 2> gpurun_out/r2c1_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c1_train_b32_bf16.json 2> gpurun_out/r2c1_train_b32_bf16.err; echo "train32 bf16 rc=$?"
HUPR_FUSED_ATTN_BWD=0 timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c1_train_b32_unfused.json 2> gpurun_out/r2c1_train_b32_unfused.err; echo "train32 unfused rc=$?"
HUPR_HALO1_OFF=1 timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c1_train_b32_bf16_nohalo.json 2> gpurun_out/r2c1_train_b32_bf16_nohalo.err; echo "train32 bf16 nohalo rc=$?"
timeout 300 python bench.py --workload forward-b1 --steps 100 --no-cpu-baseline > gpurun_out/r2c1_b1.json 2> gpurun_out/r2c1_b1.err; echo "b1 rc=$?"
timeout 300 python bench.py --single-bf16 --no-cpu-baseline > gpurun_out/r2c1_e2e_bf16.json 2> gpurun_out/r2c1_e2e_bf16.err; echo "e2e bf16 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_train_step_launches.csv python tools_dev/train_one_step.py 32 3 > gpurun_out/r2c1_ncu_train.log 2>&1; echo "train list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_train_step_bf16_launches.csv python tools_dev/train_one_step.py 32 1 > gpurun_out/r2c1_ncu_train_bf16.log 2>&1; echo "train bf16 list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -f --profile-from-start off -o gpurun_out/r02_small python tools_dev/prof_small.py > gpurun_out/r2c1_ncu_small.log 2>&1; echo "small ncu rc=$?"
python - <<'PY'
import json
for f in ("r2c1_e2e","r2c1_train_b32","r2c1_train_b32_bf16","r2c1_train_b32_unfused","r2c1_train_b32_bf16_nohalo","r2c1_b1","r2c1_e2e_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("kernels_per_replay"))
    except Exception as e:
        print(f, "ERR", e)
PY
ls -la gpurun_out | tail -20
