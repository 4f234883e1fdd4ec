#!/bin/bash
# round 2, GPU call 2: re-run of the fixed tests with programmatic dependent launch on (default) and off, PDL A/B timings, ncu --set full
# of the single-product kernels and the fused attention backward, compute-sanitizer on small cases, training bench lines.
set -u
mkdir -p gpurun_out
: > gpurun_out/r2c2_pytest.log
timeout 1200 python -m pytest tests -m gpu -q -s --durations=8 >> gpurun_out/r2c2_pytest.log 2>&1; echo "pytest (PDL on) rc=$?"
HUPR_PDL=0 timeout 600 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -m gpu -q >> gpurun_out/r2c2_pytest.log 2>&1; echo "pytest model+pipeline (PDL off) rc=$?"
grep -E "passed|failed|error" gpurun_out/r2c2_pytest.log | tail -6
for pdl in 1 0; do
  HUPR_PDL=$pdl timeout 300 python bench.py --workload forward-b1 --steps 200 --no-cpu-baseline > gpurun_out/r2c2_b1_pdl$pdl.json 2> gpurun_out/r2c2_b1_pdl$pdl.err; echo "b1 pdl=$pdl rc=$?"
  HUPR_PDL=$pdl timeout 300 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2c2_e2e_pdl$pdl.json 2> gpurun_out/r2c2_e2e_pdl$pdl.err; echo "e2e pdl=$pdl rc=$?"
done
timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c2_train_b32.json 2> gpurun_out/r2c2_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c2_train_b32_bf16.json 2> gpurun_out/r2c2_train_b32_bf16.err; echo "train32 bf16 rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo128_np1 python tools_dev/prof_kernels.py conv128_1 32 > gpurun_out/r2c2_ncu_a.log 2>&1; echo "ncu halo128 np1 rc=$?"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo64_np1 python tools_dev/prof_kernels.py conv64_1 32 > gpurun_out/r2c2_ncu_b.log 2>&1; echo "ncu halo64 np1 rc=$?"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo64_np3 python tools_dev/prof_kernels.py conv64 32 > gpurun_out/r2c2_ncu_c.log 2>&1; echo "ncu halo64 np3 rc=$?"
timeout 300 $NCU -k regex:wgrad_kernel -c 3 -o gpurun_out/r02_wgrad_np1 python tools_dev/prof_kernels.py wgrad_1 16 > gpurun_out/r2c2_ncu_d.log 2>&1; echo "ncu wgrad np1 rc=$?"
timeout 300 $NCU -k regex:attention_bwd --launch-skip 1 -c 1 -o gpurun_out/r02_attnbwd python tools_dev/prof_kernels.py attnbwd_fused 8 > gpurun_out/r2c2_ncu_e.log 2>&1; echo "ncu attnbwd rc=$?"
# compute-sanitizer: memcheck and racecheck of the mbarrier / TMEM kernels on their smallest test cases
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_conv_gemm_gpu.py -q -x -k "plain_gemm_split and 128-64-64 or test_conv_matches_torch and shape2 or cooperative_split_k and shape0" > gpurun_out/r02_sanitizer_memcheck_conv.log 2>&1; echo "memcheck conv rc=$?"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_model_gpu.py tests/test_attention_bwd_gpu.py tests/test_wgrad_gpu.py tests/test_pack_gpu.py -q -x -k "fused_attention_matches_torch and 1-128 or attention_bwd and 256 or wgrad_direct_matches_autograd and shape1 or small_layout" > gpurun_out/r02_sanitizer_memcheck_attn.log 2>&1; echo "memcheck attn/wgrad rc=$?"
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_conv_gemm_gpu.py tests/test_cascade_gpu.py -q -x -k "plain_gemm_split and 128-64-64 or test_conv_matches_torch and shape2 or matches_reference_golden" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/r02_sanitizer_memcheck_conv.log gpurun_out/r02_sanitizer_memcheck_attn.log gpurun_out/r02_sanitizer_racecheck.log
python - <<'PY'
import json
for f in ("r2c2_b1_pdl1","r2c2_b1_pdl0","r2c2_e2e_pdl1","r2c2_e2e_pdl0","r2c2_train_b32","r2c2_train_b32_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
