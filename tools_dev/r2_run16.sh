#!/bin/bash
# round 2, GPU call 15: 8-warp epilogue + the two-unit 3-tap convolution (fp16 + e4m3 cross terms): unit tests, model tests, bench A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_quant_gpu.py -x -q -s > gpurun_out/r2c16_pytest_quant.log 2>&1; echo "quant pytest rc=$?"; tail -15 gpurun_out/r2c16_pytest_quant.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py tests/test_conv_gemm_gpu.py -q -s > gpurun_out/r2c16_pytest_model.log 2>&1; echo "model pytest rc=$?"; grep -i "rel err\|passed\|failed\|error" gpurun_out/r2c16_pytest_model.log | tail -30
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c16_e2e_quant.json 2> gpurun_out/r2c16_e2e_quant.err; echo "e2e quant rc=$?"
HUPR_QUANT_FUSE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c16_e2e_quant_nofuse.json 2> gpurun_out/r2c16_e2e_quant_nofuse.err; echo "e2e quant nofuse rc=$?"
HUPR_QUANT=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c16_e2e_noquant.json 2> gpurun_out/r2c16_e2e_noquant.err; echo "e2e noquant rc=$?"
timeout 400 python bench.py --workload train --steps 10 --no-cpu-baseline > gpurun_out/r2c16_train_b32.json 2> gpurun_out/r2c16_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c16_train_b32_bf16.json 2> gpurun_out/r2c16_train_b32_bf16.err; echo "train32 bf16 rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo128_q_st256 python tools_dev/prof_kernels.py conv128_q 32 > gpurun_out/r2c16_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("r2c16_e2e_quant","r2c16_e2e_quant_nofuse","r2c16_e2e_noquant","r2c16_train_b32","r2c16_train_b32_bf16"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("kernel_ms_per_step"), r.get("executed_tensor_tflops"), r.get("two_unit_share_of_flops"))
        print("   ", {k:v for k,v in list((d.get("breakdown") or {}).items())[:6]})
    except Exception as e:
        print(f, "ERR", e); import subprocess; print(subprocess.run(["tail","-5","gpurun_out/%s.err"%f],capture_output=True,text=True).stdout)
PY
