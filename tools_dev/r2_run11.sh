#!/bin/bash
# round 2, GPU call 11: ncu --set full of the tensor-core kernels after the elect.sync change (evidence + next bottleneck)
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:attention_bwd --launch-skip 1 -c 1 -o gpurun_out/r02_attnbwd_v2_elect python tools_dev/prof_kernels.py attnbwd_fused 8 > gpurun_out/r2c11_a.log 2>&1; echo "ncu attnbwd rc=$?"
timeout 300 $NCU -k regex:attention_kernel --launch-skip 1 -c 1 -o gpurun_out/r02_attn_elect python tools_dev/prof_kernels.py attn 32 > gpurun_out/r2c11_b.log 2>&1; echo "ncu attn rc=$?"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo128_elect python tools_dev/prof_kernels.py conv128 32 > gpurun_out/r2c11_c.log 2>&1; echo "ncu halo128 rc=$?"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo64_elect python tools_dev/prof_kernels.py conv64 32 > gpurun_out/r2c11_d.log 2>&1; echo "ncu halo64 rc=$?"
timeout 300 $NCU -k regex:conv_halo --launch-skip 1 -c 1 -o gpurun_out/r02_halo128_np1_elect python tools_dev/prof_kernels.py conv128_1 32 > gpurun_out/r2c11_e.log 2>&1; echo "ncu halo128 np1 rc=$?"
timeout 300 $NCU -k regex:wgrad_kernel -c 3 -o gpurun_out/r02_wgrad_np3_elect python tools_dev/prof_kernels.py wgrad 16 > gpurun_out/r2c11_f.log 2>&1; echo "ncu wgrad np3 rc=$?"
timeout 300 $NCU -k regex:wgrad_kernel -c 3 -o gpurun_out/r02_wgrad_np1_elect python tools_dev/prof_kernels.py wgrad_1 16 > gpurun_out/r2c11_g.log 2>&1; echo "ncu wgrad np1 rc=$?"
timeout 300 $NCU -k regex:channel_sums -c 2 -o gpurun_out/r02_chsums python - > gpurun_out/r2c11_h.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from hupr_b200 import train_ops as T
from hupr_b200.ops import SplitTensor
b = 32
z = SplitTensor.from_float(torch.randn(b, 8, 64, 64, 128, device="cuda"))
g = SplitTensor.from_float(torch.randn(b, 8, 64, 64, 64, device="cuda"))
m = SplitTensor.from_float(torch.randn(b, 8, 64, 64, 64, device="cuda"))
mean = torch.zeros(64, device="cuda"); rstd = torch.ones(64, device="cuda")
s = torch.zeros(2, 64, dtype=torch.float64, device="cuda")
for _ in range(2):
    T.channel_sums(T.SUMS_BN_BWD, g, 64, s[0], s[1], b=(z, 0), mask=m, mean=mean, rstd=rstd)
torch.cuda.synchronize()
PY
echo "ncu chsums rc=$?"
