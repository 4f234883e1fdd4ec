import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hupr_b200 import ops
from hupr_b200.ops import SplitTensor
torch.manual_seed(0)
ppad, cout, cin = 13888, 128, 64
a = torch.randn(cout, ppad, device="cuda"); b = torch.randn(cin, ppad, device="cuda")
A = SplitTensor.from_float(a.view(1, 1, 1, cout, ppad)); B = SplitTensor.from_float(b.view(1, cin, ppad))
ref = (a.double() @ b.double().t())
def run(k_split, off):
    out = torch.zeros(1, 1, 1, cout, cin, device="cuda")
    t = time.time()
    try:
        ops.conv_gemm(A, ppad, B, cin, out_f32=out, k_split=k_split, w_k_off=off)
        torch.cuda.synchronize()
    except Exception as e:
        print("k_split", k_split, "off", off, "EXC", str(e)[:150], flush=True); return
    if off == 0:
        r = ref
    else:
        bs = torch.zeros_like(b)
        if off > 0: bs[:, :ppad - off] = b[:, off:]
        else: bs[:, -off:] = b[:, :ppad + off]
        r = a.double() @ bs.double().t()
    print("k_split", k_split, "off", off, "err %.3g" % float((out.view(cout, cin).double() - r).abs().max() / r.abs().max()), "%.3fs" % (time.time() - t), flush=True)
run(1, 0); run(4, 0); run(217, 0); run(1, 64); run(1, 8); run(1, 1); run(1, -1); run(8, -1191)
