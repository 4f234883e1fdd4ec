#!/bin/bash
# round 2, GPU call 8 (gpurun --gpus 2): DP parity test after the tail fix, ADC ingest test, training bench lines with the compact ingest
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_training_gpu.py tests/test_ingest_gpu.py tests/test_training_step_gpu.py -m gpu -q -s > gpurun_out/r2c8_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/r2c8_pytest.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c8_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2c8_smoke.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus 2 --workload train --steps 10 --single-bf16 --no-cpu-baseline > gpurun_out/r2c8_train_2gpu_bf16_adc.json 2> gpurun_out/r2c8_train_2gpu_bf16_adc.err; echo "train 2gpu bf16 adc rc=$?"
timeout 600 $TR --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c8_e2e_2gpu.json 2> gpurun_out/r2c8_e2e_2gpu.err; echo "e2e 2gpu rc=$?"
python - <<'PY'
import json
for f in ("r2c8_train_2gpu_bf16_adc","r2c8_e2e_2gpu"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), round(d["ms_per_step"],3), d.get("e2e"), d.get("train_probe"))
    except Exception as e:
        print(f, "ERR", e)
PY
