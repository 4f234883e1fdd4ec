"""hupr_conv_wgrad (weight gradient straight from the channels-last tensors, MN-major tensor-core operands) against autograd of
torch's float64 convolution — what loss.backward() computes for the reference's nn.Conv3d / nn.Conv2d weights
(/root/reference/tools/run.py:78; /root/reference/models/layers.py:24-32,45-63,116-123,195-210).  Tolerance: 3e-5 of max |dW|
(fp32-equivalent hi/lo arithmetic, fp32 accumulation over up to 65 536 positions, atomics across split-K slices)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [
    # n, cin, cout, d, h, w, kernel, pad
    (2, 64, 128, 4, 32, 32, (3, 3, 3), (1, 1, 1)),      # cin = 64: lane blocks pair two taps (27 atoms: odd count), BN = 128
    (1, 64, 64, 2, 16, 16, (3, 3, 3), (1, 1, 1)),       # BN = 64, W = 16 -> 4-row K blocks
    (2, 128, 64, 1, 32, 32, (1, 3, 3), (0, 1, 1)),      # cin = 128: lane blocks pair two channel blocks of one tap
    (2, 64, 64, 1, 64, 64, (1, 1, 1), (0, 0, 0)),       # a single atom
    (2, 64, 64, 8, 64, 64, (8, 1, 1), (0, 0, 0)),       # temporal merge: depth-valid
    (1, 256, 256, 2, 16, 16, (3, 3, 3), (1, 1, 1)),     # BN = 256, 108 atoms -> 27 slot groups
    (2, 320, 128, 1, 64, 64, (1, 3, 3), (0, 1, 1)),     # decoder level 1 shape (cin = 320), 128-wide W tiles
    (3, 128, 512, 1, 32, 32, (1, 1, 1), (0, 0, 0)),     # q/k/v projection: two column tiles of 256
]


@pytest.mark.parametrize("shape", SHAPES)
def test_wgrad_direct_matches_autograd(shape):
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    n, cin, cout, d, h, w, kernel, pad = shape
    torch.manual_seed(21)
    x = torch.randn(n, cin, d, h, w, device="cuda", dtype=torch.float64)
    wt = torch.randn(cout, cin, *kernel, device="cuda", dtype=torch.float64, requires_grad=True)
    y = F.conv3d(x, wt, padding=pad)
    dy = torch.randn_like(y)
    y.backward(dy)
    X = SplitTensor.from_float(x.float().permute(0, 2, 3, 4, 1).contiguous())
    DY = SplitTensor.from_float(dy.float().permute(0, 2, 3, 4, 1).contiguous())
    out = ops.conv_wgrad_direct(X, 0, cin, DY, 0, cout, kernel, pad)
    torch.cuda.synchronize()
    got = out.permute(2, 1, 0).reshape(cout, cin, *kernel).double()
    err = float((got - wt.grad).abs().max() / wt.grad.abs().max())
    assert err < 3e-5, err


def test_wgrad_direct_channel_slices_and_short_channel_rows():
    """Channel offsets into wider tensors, cin padded past the tensor's 32 channels (conv0 of the encoder), cout = 14 padded to 64
    (the head), and accumulation into a caller-provided buffer."""
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(22)
    n, d, h, w = 2, 2, 16, 16
    x = torch.randn(n, 32, d, h, w, device="cuda", dtype=torch.float64)
    wt = torch.randn(64, 32, 3, 3, 3, device="cuda", dtype=torch.float64, requires_grad=True)
    y = F.conv3d(x, wt, padding=1)
    dy = torch.randn_like(y)
    y.backward(dy)
    X = SplitTensor.from_float(x.float().permute(0, 2, 3, 4, 1).contiguous())                       # 32 channels per row
    wide = torch.randn(n, d, h, w, 192, device="cuda")
    wide[..., 64:128] = dy.float().permute(0, 2, 3, 4, 1)
    DY = SplitTensor.from_float(wide)
    out = torch.full((27, 64, 64), 0.5, dtype=torch.float32, device="cuda")
    ops.conv_wgrad_direct(X, 0, 64, DY, 64, 64, (3, 3, 3), (1, 1, 1), out=out)
    torch.cuda.synchronize()
    got = (out - 0.5)[:, :32].permute(2, 1, 0).reshape(64, 32, 3, 3, 3).double()
    assert float((got - wt.grad).abs().max() / wt.grad.abs().max()) < 3e-5
    assert float((out - 0.5)[:, 32:].abs().max()) == 0.0            # zero-filled channels contribute exactly nothing


def test_wgrad_direct_rejects_bad_arguments():
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    X = SplitTensor.empty((1, 1, 16, 16, 64), "cuda")
    DY = SplitTensor.empty((1, 1, 16, 16, 64), "cuda")
    with pytest.raises(RuntimeError):
        ops.conv_wgrad_direct(X, 0, 48, DY, 0, 64, (1, 1, 1), (0, 0, 0))          # cin not a multiple of 64
    with pytest.raises(RuntimeError):
        ops.conv_wgrad_direct(X, 0, 64, DY, 0, 64, (1, 3, 3), (0, 0, 1))          # not a 'same' convolution in H
    X24 = SplitTensor.empty((1, 1, 24, 24, 64), "cuda")
    with pytest.raises(RuntimeError):
        ops.conv_wgrad_direct(X24, 0, 64, X24, 0, 64, (1, 1, 1), (0, 0, 0))       # W = 24 does not tile 64-position K blocks


@pytest.mark.parametrize("shape", [(3, 256, 256, 256, 1024, 512), (2, 1024, 1024, 128, 512, 0), (2, 4096, 4096, 64, 320, 64)])
def test_matmul_tn_batched_matches_float64(shape):
    """out[s] += a[s]^T b[s][:, off:off+c] (the attention backward's dK = dS^T Q, dV = P^T dO): batched mode of hupr_conv_wgrad."""
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    bsz, rows, cols, c, ld, off = shape
    torch.manual_seed(23)
    a = torch.randn(bsz, rows, cols, device="cuda")
    bm = torch.randn(bsz, rows, ld, device="cuda")
    out = torch.full((bsz, cols, c), 0.25, device="cuda")
    ops.matmul_tn(SplitTensor.from_float(a), SplitTensor.from_float(bm), off, c, out)
    torch.cuda.synchronize()
    ref = torch.matmul(a.double().transpose(1, 2), bm.double()[:, :, off:off + c])
    err = float(((out.double() - 0.25) - ref).abs().max() / ref.abs().max())
    assert err < 3e-5, err


@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[1], SHAPES[5]])
def test_wgrad_single_product_mode_equals_bf16_operand_reference(shape):
    """hupr_wgrad_desc.nprod = 1 (bf16 training mode): only the hi planes are contracted; reference = float64 autograd on the
    bf16-rounded operands."""
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    n, cin, cout, d, h, w, kernel, pad = shape
    torch.manual_seed(23)
    x = torch.randn(n, cin, d, h, w, device="cuda")
    dy = torch.randn(n, cout, d + 2 * pad[0] - kernel[0] + 1, h, w, device="cuda")
    wt = torch.zeros(cout, cin, *kernel, device="cuda", dtype=torch.float64, requires_grad=True)
    F.conv3d(x.bfloat16().double(), wt, padding=pad).backward(dy.bfloat16().double())
    X = SplitTensor.from_float(x.permute(0, 2, 3, 4, 1).contiguous())
    DY = SplitTensor.from_float(dy.permute(0, 2, 3, 4, 1).contiguous())
    out = ops.conv_wgrad_direct(X, 0, cin, DY, 0, cout, kernel, pad, nprod=1)
    torch.cuda.synchronize()
    got = out.permute(2, 1, 0).reshape(cout, cin, *kernel).double()
    err = float((got - wt.grad).abs().max() / wt.grad.abs().max())
    assert err < 3e-5, err
