#!/bin/bash
# round 2, GPU call 22: measurement batch after the CTA-pair / two-unit convolution work (tests, smoke, bench lines, launch list, ncu captures)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/g2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g2_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/g2_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/g2_e2e.json 2> gpurun_out/g2_e2e.err; echo "e2e rc=$?"
HUPR_QUANT=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g2_e2e_3prod.json 2> gpurun_out/g2_e2e_3prod.err; echo "e2e 3prod rc=$?"
HUPR_QUANT=0 HUPR_HALO_SINGLE=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/g2_e2e_3prod_single.json 2> gpurun_out/g2_e2e_3prod_single.err; echo "e2e 3prod single rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/g2_ref.json 2> gpurun_out/g2_ref.err; echo "ref rc=$?"
timeout 400 python bench.py --workload train --steps 10 > gpurun_out/g2_train_b32.json 2> gpurun_out/g2_train_b32.err; echo "train32 rc=$?"
timeout 400 python bench.py --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/g2_train_b32_bf16.json 2> gpurun_out/g2_train_b32_bf16.err; echo "train32 bf16 rc=$?"
timeout 300 python bench.py --workload forward-b1 --steps 200 > gpurun_out/g2_b1.json 2> gpurun_out/g2_b1.err; echo "b1 rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_e2e_launches_v2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/g2_ncu_e2e.log 2>&1; echo "e2e list rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo128_pairs python tools_dev/prof_kernels.py conv128 32 > gpurun_out/g2_ncu_a.log 2>&1; echo "ncu pairs rc=$?"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo128_q_pairs python tools_dev/prof_kernels.py conv128_q 32 > gpurun_out/g2_ncu_b.log 2>&1; echo "ncu q pairs rc=$?"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo128_np1_pairs python tools_dev/prof_kernels.py conv128_1 32 > gpurun_out/g2_ncu_c.log 2>&1; echo "ncu np1 pairs rc=$?"
python - <<'PY'
import json
for f in ("g2_e2e","g2_e2e_3prod","g2_e2e_3prod_single","g2_ref","g2_train_b32","g2_train_b32_bf16","g2_b1"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
