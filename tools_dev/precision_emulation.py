"""CPU emulation of candidate tensor-core arithmetic for the inference forward ("spend the 37x error margin").

Every contraction of the torch-CPU oracle network (oracle/model.py) is re-run with its operands quantised the way a candidate scheme
would feed the tensor cores, and the whole-network heat maps are compared with the fp32 oracle:
    bf16x3        hi/lo bf16 split, products hi*hi + lo*hi + hi*lo            (what the library runs; validates the emulation: the
                                                                               measured GPU error is 2.7e-5)
    bf16x1        single bf16 product                                          (--single-bf16)
    fp16x1        single fp16 product (11-bit significands at the bf16 rate)
    bf16+fp8x     bf16 hi*hi at full rate + the two cross terms as e4m3 x e4m3 products (2x rate): 1 + 1/2 + 1/2 = 2 units
    fp16+fp8x     fp16 hi*hi + e4m3 cross terms
The cross terms are quantised per tensor with a power-of-two scale (what a kind::f8f6f4 MMA into the same accumulator would need).
Run here (no GPU): python tools_dev/precision_emulation.py [n_seeds]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import loss as oloss      # noqa: E402
from oracle import model as om        # noqa: E402

MODE = None
_conv3d, _conv2d, _einsum, _matmul = F.conv3d, F.conv2d, torch.einsum, torch.matmul


def hi_lo(x, dt):
    hi = x.to(dt).float()
    return hi, x - hi


def q8(x):
    """e4m3 with a per-tensor power-of-two scale that puts max|x| near 256 (inside e4m3's 448 range)."""
    m = float(x.abs().max())
    if m == 0.0:
        return x
    s = 2.0 ** torch.floor(torch.log2(torch.tensor(256.0 / m))).item()
    return (x * s).to(torch.float8_e4m3fn).float() / s


def e4m3(x, shift):
    """e4m3 of x * 2**shift with saturation at +-448 (cvt.rn.satfinite.e4m3x2.f32), returned UNscaled."""
    s = 2.0 ** shift
    return (x * s).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float() / s


def contract_halo_f16f8(fn, a, w):
    """The scheme csrc/conv_halo.cu would run for the 3x3(x3) convolutions only (every other contraction stays bf16x3):
    main product fp16(a * 2^2) x fp16(w * 2^14), cross terms e4m3(a_lo * 2^12) x e4m3(w * 2^4) and e4m3(a * 2^1) x e4m3(w_lo * 2^15),
    all three at the common scale 2^16 in ONE fp32 accumulator; fixed scales, saturating conversions."""
    a16 = (a * 4.0).clamp(-65504.0, 65504.0).to(torch.float16).float() / 4.0
    w16 = (w * 16384.0).clamp(-65504.0, 65504.0).to(torch.float16).float() / 16384.0
    a_lo, w_lo = a - a16, w - w16
    return fn(a16, w16) + fn(e4m3(a_lo, 12), e4m3(w, 4)) + fn(e4m3(a, 1), e4m3(w_lo, 15))


def contract(fn, a, b, halo=False):
    if MODE is None:
        return fn(a, b)
    if MODE == "halo-f16f8":
        if halo:
            return contract_halo_f16f8(fn, a, b)
        dt = torch.bfloat16
        a_hi, a_lo = hi_lo(a, dt)
        b_hi, b_lo = hi_lo(b, dt)
        return fn(a_hi, b_hi) + fn(a_lo.to(dt).float(), b_hi) + fn(a_hi, b_lo.to(dt).float())
    if MODE in ("bf16x1", "fp16x1"):
        dt = torch.bfloat16 if MODE == "bf16x1" else torch.float16
        return fn(a.to(dt).float(), b.to(dt).float())
    dt = torch.float16 if MODE.startswith("fp16") else torch.bfloat16
    a_hi, a_lo = hi_lo(a, dt)
    b_hi, b_lo = hi_lo(b, dt)
    if MODE.endswith("x3"):
        a_lo, b_lo = a_lo.to(dt).float(), b_lo.to(dt).float()
        return fn(a_hi, b_hi) + fn(a_lo, b_hi) + fn(a_hi, b_lo)
    return fn(a_hi, b_hi) + fn(q8(a_lo), q8(b_hi)) + fn(q8(a_hi), q8(b_lo))      # "+fp8x"


def install():
    F.conv3d = lambda x, w, bias=None, **k: contract(lambda a, b: _conv3d(a, b, **k), x, w, halo=w.shape[-2] == 3) + (0 if bias is None else bias.view(1, -1, 1, 1, 1))
    F.conv2d = lambda x, w, bias=None, **k: contract(lambda a, b: _conv2d(a, b, **k), x, w, halo=w.shape[-2] == 3) + (0 if bias is None else bias.view(1, -1, 1, 1))
    om.torch.einsum = lambda eq, a, b: contract(lambda p, q: _einsum(eq, p, q), a, b)
    om.torch.matmul = lambda a, b: contract(_matmul, a, b)


MODES = ("bf16x3", "bf16x1", "fp16x1", "bf16+fp8x", "fp16+fp8x", "halo-f16f8")


def main(n_seeds):
    global MODE
    install()
    rows = {}
    for seed in range(n_seeds):
        sd = om.make_state_dict(seed)
        h, v = om.make_vrdae(1, seed)
        MODE = None
        with torch.no_grad():
            rh, rg = om.huprnet_forward(sd, h, v)
        rk, _ = oloss.get_max_preds(rg.view(1, 14, 64, 64).numpy())
        for m in MODES:
            MODE = m
            with torch.no_grad():
                gh, gg = om.huprnet_forward(sd, h, v)
            e = max(float(((gh - rh).abs() / rh.abs()).max()), float(((gg - rg).abs() / rg.abs()).max()))
            k, _ = oloss.get_max_preds(gg.view(1, 14, 64, 64).numpy())
            rows.setdefault(m, []).append((e, int((k != rk).any(-1).sum())))
            print("seed %d %-10s max element-wise relative heat-map error %.2e, keypoints moved %d/14" % (seed, m, e, rows[m][-1][1]), flush=True)
    print("\nworst case over %d seeds (north_star tolerance 1e-3):" % n_seeds)
    for m, r in rows.items():
        print("  %-10s %.2e   keypoints moved %d/%d" % (m, max(e for e, _ in r), sum(k for _, k in r), 14 * len(r)))


if __name__ == "__main__":
    if len(sys.argv) > 2:
        MODES = tuple(sys.argv[2].split(","))
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
