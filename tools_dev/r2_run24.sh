#!/bin/bash
# round 2, GPU call 24: interleaved e4m3 operand plane (64-byte rows) for the two-unit convolutions
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_quant_gpu.py -x -q -s > gpurun_out/r2c24_pytest_quant.log 2>&1; echo "quant pytest rc=$?"; tail -8 gpurun_out/r2c24_pytest_quant.log
timeout 120 python tools_dev/time_halo_grid.py > gpurun_out/r2c24_halo_grid.txt 2>&1; echo "grid sweep rc=$?"; grep "mode q\|mode 3 grid 148" gpurun_out/r2c24_halo_grid.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c24_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c24_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c24_e2e.json 2> gpurun_out/r2c24_e2e.err; echo "e2e rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo128_q64_pairs python tools_dev/prof_kernels.py conv128_q 32 > gpurun_out/r2c24_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("r2c24_e2e",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(f, round(d["value"],1), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("kernel_ms_per_step"), d.get("clocks"))
        for k,v in list(d["breakdown"]["conv_gemm_by_shape"].items())[:5]: print("      ",k,v)
    except Exception as e:
        print(f, "ERR", e)
PY
