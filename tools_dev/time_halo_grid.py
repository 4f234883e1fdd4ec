"""Is the halo convolution bound per SM or chip-wide?  Times one 64 -> 128 channel 3x3x3 layer (batch 32) in its three arithmetic modes with
148, 74 and 37 persistent CTAs (HUPR_HALO_GRID): if the time per tile of a CTA does not change with the number of CTAs the limit is per SM
(TMA / shared-memory side); if fewer CTAs each run faster, they compete for a chip-wide resource (L2 -> SM operand bandwidth)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hupr_b200 import ops
from hupr_b200.ops import SplitTensor

b = 32
torch.manual_seed(0)
x = SplitTensor.from_float(torch.randn(b, 8, 64, 64, 64, device="cuda"))
w = SplitTensor.from_float(torch.randn(27, 128, 64, device="cuda") * 0.02)
out = SplitTensor.empty((b, 8, 64, 64, 128), "cuda")
tiles = b * 8 * 16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(mode):
    if mode == "q":
        with ops.quant():
            ops.conv_gemm(x, 64, w, 128, kernel=(3, 3, 3), pad=(1, 1, 1), out=out)
    elif mode == "1":
        with ops.products(1):
            ops.conv_gemm(x, 64, w, 128, kernel=(3, 3, 3), pad=(1, 1, 1), out=out)
    else:
        ops.conv_gemm(x, 64, w, 128, kernel=(3, 3, 3), pad=(1, 1, 1), out=out)


with ops.quant():
    ops.quantize_planes(x)
    x.q_fresh = (0, 64)
for mode in ("3", "q", "1"):
    for grid in (148, 74, 37):
        os.environ["HUPR_HALO_GRID"] = str(grid)
        ts = []
        for it in range(6):
            flush.zero_()
            if mode == "q":
                x.q_fresh = (0, 64)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); run(mode); e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ms = sorted(ts[1:])[len(ts[1:]) // 2]
        per_cta = tiles / grid
        print("mode %s grid %3d: %.3f ms, %.1f tiles per CTA, %.2f us per tile per CTA, chip %.1f tiles/us" % (mode, grid, ms, per_cta, ms * 1e3 / per_cta, tiles / (ms * 1e3)), flush=True)
