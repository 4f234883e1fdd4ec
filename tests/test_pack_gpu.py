"""GPU tests of the parameter-layout kernels (csrc/pack.cu): they replace torch's permute / flip / cat / copy_ element-wise kernels
inside the training step, so each is compared with exactly that torch expression (bit-exact: same round-to-nearest-even bf16 split)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("cout,cin,kernel,cout_off,cout_total", [
    (64, 32, (3, 3, 3), 0, 64),          # encoder layer1.0: cin padded 32 -> 64
    (128, 128, (3, 3, 3), 128, 256),     # second half of a concatenated main || downsample filter
    (14, 32, (1, 1, 1), 0, 64),          # head: 14 keypoints in a 64-wide tile
    (32, 64, (1, 3, 3), 64, 128),        # decoderLayer1.1: cout 32 padded to 64, second slot
    (256, 256, (2, 1, 1), 0, 256),       # temporal merge
    (1024, 1024, (1, 1, 1), 0, 1024),    # GCN weight: forward operand = W, data-gradient operand = W^T
])
def test_pack_conv_weights_matches_torch_packing(cout, cin, kernel, cout_off, cout_total):
    from hupr_b200 import ops
    from hupr_b200.models import layers as L
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(cout + cin)
    w = torch.randn((cout, cin) + kernel, device=DEV) * 0.1
    taps = kernel[0] * kernel[1] * kernel[2]
    cin_pad, cpad = L.pad64(cin), L.pad64(cout)
    fwd = SplitTensor.empty((taps, cout_total, cin_pad), DEV, True, zero=True)
    dgr = SplitTensor.empty((taps, cin_pad, cout_total), DEV, True, zero=True)
    ops.pack_conv_weights(w, fwd, cout_off, dgr)
    torch.cuda.synchronize()
    ref_f = L.pack_conv([w], cin_pad, [cpad])                                   # [taps, cpad, cin_pad]
    full = torch.zeros((cpad, cin_pad) + kernel, device=DEV)
    full[:cout, :cin] = w
    ref_d = L.pack_dgrad(full, cpad, cin_pad)                                   # [taps, cin_pad, cpad]
    for got, ref in ((fwd.hi, ref_f.hi), (fwd.lo, ref_f.lo)):
        assert torch.equal(got[:, cout_off:cout_off + cpad], ref)
    for got, ref in ((dgr.hi, ref_d.hi), (dgr.lo, ref_d.lo)):
        assert torch.equal(got[:, :, cout_off:cout_off + cpad], ref)
    # nothing outside the slot was touched
    mask = torch.ones(cout_total, dtype=torch.bool, device=DEV)
    mask[cout_off:cout_off + cpad] = False
    assert float(fwd.hi[:, mask].float().abs().sum()) == 0 and float(dgr.hi[:, :, mask].float().abs().sum()) == 0


@pytest.mark.parametrize("cout,cin,taps,cout_off,cout_total", [(64, 32, 27, 0, 64), (128, 128, 27, 128, 256), (14, 32, 1, 0, 64), (32, 64, 9, 64, 128)])
def test_unpack_wgrad_matches_permute(cout, cin, taps, cout_off, cout_total):
    from hupr_b200 import ops
    from hupr_b200.models import layers as L
    torch.manual_seed(taps)
    cin_pad = L.pad64(cin)
    acc = torch.randn((taps, cin_pad, cout_total), device=DEV)
    dst = torch.full((cout, cin, taps), float("nan"), device=DEV)
    ops.unpack_wgrad(acc, cout_off, dst)
    torch.cuda.synchronize()
    assert torch.equal(dst, acc[:, :cin, cout_off:cout_off + cout].permute(2, 1, 0).contiguous())


def test_small_layout_kernels():
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(0)
    src = torch.randn(5 * 64, dtype=torch.float64, device=DEV)
    out = torch.empty(5, dtype=torch.float32, device=DEV)
    ops.reduce_f64(src, 5, 64, out)
    assert torch.allclose(out.double(), src.view(5, 64).sum(1), rtol=1e-6, atol=1e-6)
    cast = torch.empty(320, dtype=torch.float32, device=DEV)
    ops.reduce_f64(src, 320, 1, cast)
    assert torch.equal(cast, src.float())
    slope = torch.tensor([0.25], device=DEV)
    arr = torch.empty(192, device=DEV)
    ops.broadcast_f32(slope, arr)
    assert torch.equal(arr, slope.expand(192))
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    for _ in range(3):
        ops.bump_i32(counter)
    assert int(counter) == 3
    z = torch.ones(1000, device=DEV)
    ops.zero_(z)
    assert float(z.abs().sum()) == 0
    # GCN bias rows: rows[(b, j)][q] = bias[q][j], zero rows past batch * 14
    batch, rows = 3, 128
    bias = torch.randn(1024, 14, device=DEV)
    got = SplitTensor.empty((1, 1, 1, rows, 1024), DEV, True)
    got.hi.fill_(7.0); got.lo.fill_(7.0)
    ops.gcn_bias_rows(bias, batch, got)
    t = torch.zeros(rows, 1024, device=DEV)
    t[:batch * 14] = bias.t().repeat(batch, 1)
    ref = SplitTensor.from_float(t.view(1, 1, 1, rows, 1024))
    torch.cuda.synchronize()
    assert torch.equal(got.hi, ref.hi) and torch.equal(got.lo, ref.lo)


def test_zero_arena_replays_the_planned_sequence():
    from hupr_b200.training import ZeroArena
    arena = ZeroArena(torch.device(DEV))
    shapes = [((2, 64), torch.float64), ((27, 64, 128), torch.float32), ((5,), torch.float64)]
    for _ in range(3):
        arena.begin()
        ts = [arena.take(s, d) for s, d in shapes]
        for t, (s, d) in zip(ts, shapes):
            assert tuple(t.shape) == s and t.dtype == d and float(t.double().abs().sum()) == 0
            t.fill_(3.0)                       # dirtied: the next begin() must clear it again
    assert arena.buf is not None and arena.buf.numel() == sum(arena.sizes)
    arena.begin()
    arena.take((2, 64), torch.float64)
    with pytest.raises(RuntimeError, match="sequence changed"):
        arena.take((3,), torch.float32)
