"""Where is the floor of the [S,S]-output GEMMs?  Times a 1 GiB device fill / copy and the two attention-backward GEMMs alone."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hupr_b200 import ops
from hupr_b200.ops import SplitTensor


def timeit(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


b, s_, c = 16, 4096, 64
n = b * s_ * s_
x = torch.empty(n, dtype=torch.bfloat16, device="cuda"); y = torch.empty(n, dtype=torch.bfloat16, device="cuda")
x2 = torch.empty(n, dtype=torch.bfloat16, device="cuda")
t = timeit(lambda: (x.fill_(1.0), x2.fill_(1.0)))
print("fill 2 x %.2f GB: %.3f ms = %.0f GB/s written" % (n * 2 / 1e9, t, 2 * n * 2 / t / 1e6))
t = timeit(lambda: y.copy_(x))
print("copy %.2f GB: %.3f ms = %.0f GB/s read+write" % (n * 2 / 1e9, t, 2 * n * 2 / t / 1e6))
q = SplitTensor.from_float(torch.randn(b, 1, 1, s_, c, device="cuda") * 0.5)
k = SplitTensor.from_float(torch.randn(b, s_, c, device="cuda") * 0.5)
lse = torch.full((b, s_), 8.0, device="cuda")
probs = SplitTensor.empty((b, 1, 1, s_, s_), "cuda")
ds = SplitTensor.empty((b, 1, 1, s_, s_), "cuda")
scratch = torch.empty((b, 1, 1, s_, s_), dtype=torch.float32, device="cuda")
for name, fn in (
        ("P = exp(QK^T - lse)   [TMA store]", lambda: ops.conv_gemm(q, c, k, s_, w_batched=True, out=probs, row_vec=lse, row_mode=1)),
        ("P = exp(QK^T - lse)   [row stores]", lambda: ops.conv_gemm(q, c, k, s_, w_batched=True, out=probs, row_vec=lse, row_mode=1, tma_store=False)),
        ("plain QK^T -> hi/lo   [TMA store]", lambda: ops.conv_gemm(q, c, k, s_, w_batched=True, out=probs)),
        ("plain QK^T -> fp32", lambda: ops.conv_gemm(q, c, k, s_, w_batched=True, out_f32=scratch)),
        ("dS = P*(dOV^T - rd)   [TMA store]", lambda: ops.conv_gemm(q, c, k, s_, w_batched=True, out=ds, residual=probs, row_vec=lse, row_mode=2)),
):
    t = timeit(fn, 5)
    print("%-36s %.3f ms  (%.0f GB/s of output)" % (name, t, n * 4 / t / 1e6))
