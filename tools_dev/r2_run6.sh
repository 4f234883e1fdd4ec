#!/bin/bash
# round 2, GPU call 6: fused BatchNorm statistics — tests, A/B training bench lines, ncu of wgrad after the rework and channel_sums
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2c6_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/r2c6_pytest.log | tail -4
timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c6_train_b32.json 2> gpurun_out/r2c6_train_b32.err; echo "train32 rc=$?"
HUPR_FUSED_BN_STATS=0 timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c6_train_b32_nofuse.json 2> gpurun_out/r2c6_train_b32_nofuse.err; echo "train32 nofuse rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c6_train_b32_bf16.json 2> gpurun_out/r2c6_train_b32_bf16.err; echo "train32 bf16 rc=$?"
timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2c6_e2e.json 2> gpurun_out/r2c6_e2e.err; echo "e2e rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:wgrad_kernel -c 3 -o gpurun_out/r02_wgrad_np3_v2 python tools_dev/prof_kernels.py wgrad 16 > gpurun_out/r2c6_ncu_a.log 2>&1; echo "ncu wgrad np3 rc=$?"
timeout 300 $NCU -k regex:wgrad_kernel -c 3 -o gpurun_out/r02_wgrad_np1_v2 python tools_dev/prof_kernels.py wgrad_1 16 > gpurun_out/r2c6_ncu_b.log 2>&1; echo "ncu wgrad np1 rc=$?"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo64_ncat python tools_dev/prof_kernels.py conv64 32 > gpurun_out/r2c6_ncu_c.log 2>&1; echo "ncu halo64 ncat rc=$?"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo128_np1_v2 python tools_dev/prof_kernels.py conv128_1 32 > gpurun_out/r2c6_ncu_d.log 2>&1; echo "ncu halo128 np1 rc=$?"
python - <<'PY'
import json
for f in ("r2c6_e2e","r2c6_train_b32","r2c6_train_b32_nofuse","r2c6_train_b32_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("launches_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
