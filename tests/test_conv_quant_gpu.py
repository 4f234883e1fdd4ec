"""GPU tests of the two-unit 3-tap convolution (hupr_conv_desc.nprod == 2, csrc/conv_halo.cu NPROD == 2): fp16 main product + the two
hi*lo cross terms as e4m3 products, one fp32 accumulator.  Replaces the fp32 contraction of nn.Conv3d / nn.Conv2d
(/root/reference/models/layers.py:24-32,45-63) for the large eval-mode layers.

Checks: (1) the operand planes (hupr_quantize_planes) against the same roundings done by torch; (2) the convolution against a torch
emulation of exactly this arithmetic (tight: only fp32 summation order differs) and against the fp64 convolution (the accuracy the scheme
is supposed to have: a few 1e-5 of the output RMS); (3) the planes a producing convolution writes from its epilogue; (4) that shapes the
kernel does not take fall back to the three bf16 products bit-exactly."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _e4m3(x, scale):
    return (x * scale).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float() / scale


def _unpack_q8(q8, shape):
    """e4m3 plane [..., C/32, 2, 32] bytes -> (values, residuals) as float32 tensors of ``shape`` (still scaled)."""
    v = q8.view(torch.float8_e4m3fn).float().view(tuple(shape[:-1]) + (shape[-1] // 32, 2, 32))
    return v[..., 0, :].reshape(shape), v[..., 1, :].reshape(shape)


def _planes(x, is_weight):
    """(x16, x8, x8l) as UNscaled float32 values, rounded the way hupr_quantize_planes rounds them."""
    s16, s8, s8l = (16384.0, 16.0, 32768.0) if is_weight else (4.0, 2.0, 4096.0)
    x16 = (x * s16).clamp(-65504.0, 65504.0).to(torch.float16).float() / s16
    return x16, _e4m3(x, s8), _e4m3(x - x16, s8l)


@pytest.mark.parametrize("is_weight", [False, True])
def test_quantize_planes_match_torch_roundings(is_weight):
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(3)
    x = torch.randn(4, 2, 8, 16, 64, device=DEV) * (0.05 if is_weight else 3.0)
    x[0, 0, 0, 0, :8] = torch.tensor([0.0, 1e-6, -1e-6, 300.0, -300.0, 5000.0, 1e-3, -2.5], device=DEV) * (0.01 if is_weight else 1.0)
    t = SplitTensor.from_float(x)
    ops.quantize_planes(t, 0, None, is_weight)
    # a channel sub-range leaves the other channels untouched
    u = SplitTensor.from_float(x)
    for p in u.ensure_q():
        p.zero_()
    ops.quantize_planes(u, 32, 32, is_weight)
    torch.cuda.synchronize()
    s16, s8, s8l = (16384.0, 16.0, 32768.0) if is_weight else (4.0, 2.0, 4096.0)
    ref16, ref8, ref8l = _planes(t.float(), is_weight)
    q16, q8x = t.q
    q8, q8l = _unpack_q8(q8x, x.shape)
    assert torch.equal(q16.float() / s16, ref16)
    assert torch.equal(q8 / s8, ref8)
    assert torch.equal(q8l / s8l, ref8l)
    u8, u8l = _unpack_q8(u.q[1], x.shape)
    assert torch.equal(u.q[0][..., 32:], q16[..., 32:]) and float(u.q[0][..., :32].float().abs().sum()) == 0
    assert torch.equal(u8l[..., 32:], q8l[..., 32:]) and torch.equal(u8[..., 32:], q8[..., 32:]) and int(u.q[1][..., :64].sum()) == 0
    # the planes reproduce the value: x16 + x8l is within 2^-11 (fp16 residual) * 2^-4 (e4m3 rounding) of |x| wherever the residual is in
    # e4m3's normal range; small values lose part of the correction (flush to e4m3 subnormals), never more than the fp16 residual itself
    val = t.float()
    rec = ref16 + ref8l
    big = (val.abs() > (1e-3 if is_weight else 0.25)) & (val.abs() < (0.05 if is_weight else 200.0))
    assert float(((rec - val).abs() / val.abs())[big].max()) < 2.0 ** -14
    mid = (val.abs() > (1e-6 if is_weight else 1e-4)) & (val.abs() < (0.05 if is_weight else 200.0))      # above fp16's subnormal range after scaling
    assert float(((rec - val).abs() / val.abs())[mid].max()) <= 2.0 ** -11


def _reference_conv(x, w, pad):
    """x [N, D, H, W, C] and w [taps, cout, cin] float64 -> [N, D', H, W, cout] float64 (same 'same'-padded correlation as the kernel)."""
    kd = w.shape[0] // 9
    wt = w.view(kd, 3, 3, w.shape[1], w.shape[2]).permute(3, 4, 0, 1, 2)
    return F.conv3d(x.permute(0, 4, 1, 2, 3), wt, padding=pad).permute(0, 2, 3, 4, 1)


@pytest.mark.parametrize("n,d,hw,cin,cout,kd", [(16, 2, 32, 64, 128, 3), (4, 1, 64, 128, 256, 1), (64, 2, 16, 128, 128, 3)])
def test_two_unit_conv_matches_its_emulation_and_the_fp64_convolution(n, d, hw, cin, cout, kd):
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(n + cin)
    x = torch.randn(n, d, hw, hw, cin, device=DEV) * 2.0
    w = torch.randn(kd * 9, cout, cin, device=DEV) * (2.0 / (kd * 9 * cin)) ** 0.5
    xs, ws = SplitTensor.from_float(x), SplitTensor.from_float(w)
    pad = (kd // 2, 1, 1)
    scale = torch.rand(cout, device=DEV) + 0.5
    shift = torch.randn(cout, device=DEV)
    out3 = SplitTensor.empty((n, d, hw, hw, cout), DEV)
    ops.conv_gemm(xs, cin, ws, cout, kernel=(kd, 3, 3), pad=pad, scale=scale, shift=shift, out=out3)
    out2 = SplitTensor.empty((n, d, hw, hw, cout), DEV)
    with ops.quant():
        assert ops.conv_gemm(xs, cin, ws, cout, kernel=(kd, 3, 3), pad=pad, out=out2, probe=True)
        ops.conv_gemm(xs, cin, ws, cout, kernel=(kd, 3, 3), pad=pad, scale=scale, shift=shift, out=out2, out_q=True)
    torch.cuda.synchronize()
    xv, wv = xs.float().double(), ws.float().double()
    ref = _reference_conv(xv, wv, pad) * scale.double() + shift.double()
    x16, x8, x8l = (t.double() for t in _planes(xs.float(), False))
    w16, w8, w8l = (t.double() for t in _planes(ws.float(), True))
    emu = (_reference_conv(x16, w16, pad) + _reference_conv(x8l, w8, pad) + _reference_conv(x8, w8l, pad)) * scale.double() + shift.double()
    top = float(ref.abs().max())         # the bound of the other convolution tests: 2e-5 of the tensor's max-abs (hi + lo bf16 storage of the output
    e_emu = float((out2.float().double() - emu).abs().max()) / top      # carries ~2^-17 of a value, fp32 accumulation order the rest)
    e_ref2 = float((out2.float().double() - ref).abs().max()) / top
    e_ref3 = float((out3.float().double() - ref).abs().max()) / top
    print("two-unit conv %s: max error / max |output| vs its emulation %.3g, vs fp64 %.3g (three bf16 products: %.3g)" % ((n, d, hw, cin, cout, kd), e_emu, e_ref2, e_ref3))
    assert e_emu < 2e-5          # the kernel computes exactly this arithmetic
    assert e_ref2 < 4e-5         # the accuracy of the scheme itself: ~2^-13 per product, averaging over the contraction
    assert e_ref3 < 2e-5
    # planes written by the epilogue == planes of the stored hi/lo output, up to the 2^-17 the hi/lo rounding of the output moves a value
    q16, q8x = out2.q
    q8, q8l = _unpack_q8(q8x, out2.hi.shape)
    assert out2.q_fresh == (0, cout)
    r16, r8, r8l = _planes(out2.float(), False)
    got = q16.float() / 4.0 + q8l / 4096.0
    val = out2.float()
    assert float(((got - val).abs() / val.abs().clamp_min(1e-2)).max()) < 2.0 ** -14
    assert float((q16.float() / 4.0 - r16).abs().max()) <= float(val.abs().max()) * 2.0 ** -10       # at most one fp16 ulp apart
    assert float((q8 / 2.0 - r8).abs().max()) <= float(val.abs().max()) * 2.0 ** -3
    # a second two-unit convolution consumes them without a separate pass (q_fresh is one-shot)
    w2 = SplitTensor.from_float(torch.randn(9, 256, cout, device=DEV) * (2.0 / (9 * cout)) ** 0.5)
    o_a, o_b = SplitTensor.empty((n, d, hw, hw, 256), DEV), SplitTensor.empty((n, d, hw, hw, 256), DEV)
    with ops.quant():
        before = ops.launch_count()
        ops.conv_gemm(out2, cout, w2, 256, kernel=(1, 3, 3), pad=(0, 1, 1), out=o_a)
        first = ops.launch_count() - before          # weight planes (once) + convolution, no activation pass
        assert out2.q_fresh is None
        before = ops.launch_count()
        ops.conv_gemm(out2, cout, w2, 256, kernel=(1, 3, 3), pad=(0, 1, 1), out=o_b)
        second = ops.launch_count() - before         # activation pass + convolution
    torch.cuda.synchronize()
    assert (first, second) == (2, 2)
    ref2 = _reference_conv(out2.float().double(), w2.float().double(), (0, 1, 1))
    top2 = float(ref2.abs().max())
    assert float((o_a.float().double() - ref2).abs().max()) / top2 < 4e-5
    assert float((o_b.float().double() - ref2).abs().max()) / top2 < 4e-5


def test_shapes_outside_the_two_unit_kernel_fall_back_bit_exactly():
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(0)
    for n, d, hw, cin, cout, kernel, pad in ((2, 1, 16, 64, 128, (1, 3, 3), (0, 1, 1)),      # too few tiles
                                             (16, 2, 32, 64, 64, (3, 3, 3), (1, 1, 1)),      # cout = 64 tiles keep the [w_hi | w_lo] form
                                             (8, 1, 32, 64, 128, (1, 1, 1), (0, 0, 0))):     # not a 3-tap convolution
        x = SplitTensor.from_float(torch.randn(n, d, hw, hw, cin, device=DEV))
        w = SplitTensor.from_float(torch.randn(kernel[0] * kernel[1] * kernel[2], cout, cin, device=DEV) * 0.05)
        a, b = SplitTensor.empty((n, d, hw, hw, cout), DEV), SplitTensor.empty((n, d, hw, hw, cout), DEV)
        ops.conv_gemm(x, cin, w, cout, kernel=kernel, pad=pad, out=a)
        with ops.quant():
            assert not ops.conv_gemm(x, cin, w, cout, kernel=kernel, pad=pad, out=b, probe=True)
            ops.conv_gemm(x, cin, w, cout, kernel=kernel, pad=pad, out=b)
        torch.cuda.synchronize()
        assert torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo)
        assert x.q is None and w.q is None


def test_large_activations_degrade_gracefully_and_large_weights_fall_back():
    """Beyond |a| = 224 the e4m3 operands saturate: a term then keeps single-fp16-product accuracy (2^-12) instead of ~2^-16 — the result
    stays close, nothing blows up; a filter with a weight >= 3.99 (outside the fp16 plane w * 2^14) keeps the three bf16 products."""
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(5)
    n, d, hw, cin, cout = 16, 2, 32, 64, 128
    x = torch.randn(n, d, hw, hw, cin, device=DEV) * 300.0           # |x| up to ~1500: a8 saturates for most of the tensor
    w = torch.randn(27, cout, cin, device=DEV) * (2.0 / (27 * cin)) ** 0.5
    xs, ws = SplitTensor.from_float(x), SplitTensor.from_float(w)
    out = SplitTensor.empty((n, d, hw, hw, cout), DEV)
    with ops.quant():
        ops.conv_gemm(xs, cin, ws, cout, kernel=(3, 3, 3), pad=(1, 1, 1), out=out)
    torch.cuda.synchronize()
    ref = _reference_conv(xs.float().double(), ws.float().double(), (1, 1, 1))
    err = float((out.float().double() - ref).abs().max()) / float(ref.abs().max())
    print("two-unit conv with |x| up to %.0f: max error / max |output| %.3g" % (float(x.abs().max()), err))
    assert torch.isfinite(out.float()).all() and err < 4e-4          # CPU emulation of this case: 1.3e-4 (1.2e-5 inside the range)
    assert ops.quant_saturations("cuda") == 0                        # ... and nothing left the fp16 plane (|x| < 16 376)
    huge = SplitTensor.from_float(x * 40.0)                          # |x| up to ~60 000: the fp16 plane itself saturates -> reported
    with ops.quant():
        ops.conv_gemm(huge, cin, ws, cout, kernel=(3, 3, 3), pad=(1, 1, 1), out=out, out_q=True)
    n_sat = ops.quant_saturations("cuda")
    assert n_sat > 0 and ops.quant_saturations("cuda") == 0          # counted once, reset by the read
    big = w.clone()
    big[3, 5, 7] = 4.5
    wb = SplitTensor.from_float(big)
    xs2 = SplitTensor.from_float(torch.randn(n, d, hw, hw, cin, device=DEV))
    a, b = SplitTensor.empty((n, d, hw, hw, cout), DEV), SplitTensor.empty((n, d, hw, hw, cout), DEV)
    ops.conv_gemm(xs2, cin, wb, cout, kernel=(3, 3, 3), pad=(1, 1, 1), out=a)
    with ops.quant():
        ops.conv_gemm(xs2, cin, wb, cout, kernel=(3, 3, 3), pad=(1, 1, 1), out=b)
    torch.cuda.synchronize()
    assert wb.q is False and torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo)
