"""Small driver for ncu captures of the two tensor-core kernels (run under `ncu --set full -k regex:... -c 1`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hupr_b200 import ops
from hupr_b200.ops import SplitTensor

which = sys.argv[1] if len(sys.argv) > 1 else "all"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.manual_seed(0)
if which in ("conv128", "all"):
    x = SplitTensor.from_float(torch.randn(b, 8, 64, 64, 64, device="cuda"))
    w = SplitTensor.from_float(torch.randn(27, 128, 64, device="cuda") * 0.02)
    out = SplitTensor.empty((b, 8, 64, 64, 128), "cuda")
    for _ in range(3):
        ops.conv_gemm(x, 64, w, 128, kernel=(3, 3, 3), pad=(1, 1, 1), out=out)
if which in ("conv64", "all"):
    x = SplitTensor.from_float(torch.randn(b, 8, 64, 64, 128, device="cuda"))
    w = SplitTensor.from_float(torch.randn(27, 64, 64, device="cuda") * 0.02)
    out = SplitTensor.empty((b, 8, 64, 64, 64), "cuda")
    for _ in range(3):
        ops.conv_gemm(x, 64, w, 64, kernel=(3, 3, 3), pad=(1, 1, 1), residual=x, r_ch_off=64, out=out)
if which in ("attn", "all"):
    s, c = 4096, 64
    pq = SplitTensor.from_float(torch.randn(b, 1, 1, s, 4 * c, device="cuda"))
    v = SplitTensor.from_float(torch.randn(b, 1, 1, s, c, device="cuda"))
    vt = SplitTensor.empty((b, c, s), "cuda")
    ops.transpose_split(v, c, vt)
    out = SplitTensor.empty((b, 1, 1, s, 4 * c), "cuda")
    for _ in range(3):
        ops.attention_fwd(pq, c, pq, 0, vt, c, out, 0, residual=v)
if which in ("wgrad", "all"):
    # three template instantiations on training-step shapes: BN = 64 / 128 (level 1, W = 64) and BN = 256 (level 2)
    for cin, cout, d, h, w in ((64, 64, 8, 64, 64), (64, 128, 8, 64, 64), (128, 256, 4, 32, 32)):
        x = SplitTensor.from_float(torch.randn(b, d, h, w, cin, device="cuda"))
        dy = SplitTensor.from_float(torch.randn(b, d, h, w, cout, device="cuda"))
        out = torch.zeros((27, cin, cout), device="cuda")
        ops.conv_wgrad_direct(x, 0, cin, dy, 0, cout, (3, 3, 3), (1, 1, 1), out=out)
if which in ("conv128_1", "conv64_1"):
    # single-product (bf16 training mode) 3x3x3 convolutions through the halo kernel
    cout = 128 if which == "conv128_1" else 64
    x = SplitTensor.from_float(torch.randn(b, 8, 64, 64, 64, device="cuda"))
    w = SplitTensor.from_float(torch.randn(27, cout, 64, device="cuda") * 0.02)
    out = SplitTensor.empty((b, 8, 64, 64, cout), "cuda")
    with ops.products(1):
        for _ in range(3):
            ops.conv_gemm(x, 64, w, cout, kernel=(3, 3, 3), pad=(1, 1, 1), out=out)
if which in ("wgrad_1",):
    with ops.products(1):
        for cin, cout, d, h, w in ((64, 64, 8, 64, 64), (64, 128, 8, 64, 64), (128, 256, 4, 32, 32)):
            x = SplitTensor.from_float(torch.randn(b, d, h, w, cin, device="cuda"))
            dy = SplitTensor.from_float(torch.randn(b, d, h, w, cout, device="cuda"))
            out = torch.zeros((27, cin, cout), device="cuda")
            ops.conv_wgrad_direct(x, 0, cin, dy, 0, cout, (3, 3, 3), (1, 1, 1), out=out)
if which in ("attnbwd_fused",):
    s_, c = 4096, 64
    mk = lambda: SplitTensor.from_float(torch.randn(b, 1, 1, s_, 4 * c, device="cuda") * 0.3)
    pq, v, do = mk(), mk(), mk()
    lse = torch.full((b, s_), 9.0, device="cuda")
    rowdot = torch.zeros((b, s_), device="cuda")
    dq = torch.zeros((b, s_, 4 * c), device="cuda")
    for _ in range(2):
        ops.attention_bwd(pq, c, pq, 0, v, 0, do, 0, lse, rowdot, dq, c, dq, 0, dq, 2 * c)
if which in ("attnbwd",):
    # the [S, S]-output GEMMs of the attention backward at level 1: P = exp(Q K^T - lse) and dS = P * (dO V^T - rowdot)
    s_, c = 4096, 64
    q = SplitTensor.from_float(torch.randn(b, 1, 1, s_, c, device="cuda") * 0.5)
    k = SplitTensor.from_float(torch.randn(b, s_, c, device="cuda") * 0.5)
    lse = torch.full((b, s_), 8.0, device="cuda")
    probs = SplitTensor.empty((b, 1, 1, s_, s_), "cuda")
    ds = SplitTensor.empty((b, 1, 1, s_, s_), "cuda")
    for _ in range(2):
        ops.conv_gemm(q, c, k, s_, w_batched=True, out=probs, row_vec=lse, row_mode=1)
        ops.conv_gemm(q, c, k, s_, w_batched=True, out=ds, residual=probs, row_vec=lse, row_mode=2)
torch.cuda.synchronize()
print("done")
if which in ("conv128_q",):
    # two-unit arithmetic (fp16 main product + e4m3 cross terms) on the same 64 -> 128 channel 3x3x3 layer as "conv128"
    x = SplitTensor.from_float(torch.randn(b, 8, 64, 64, 64, device="cuda"))
    w = SplitTensor.from_float(torch.randn(27, 128, 64, device="cuda") * 0.02)
    out = SplitTensor.empty((b, 8, 64, 64, 128), "cuda")
    with ops.quant():
        for _ in range(3):
            ops.conv_gemm(x, 64, w, 128, kernel=(3, 3, 3), pad=(1, 1, 1), out=out, out_q=True)
torch.cuda.synchronize()
