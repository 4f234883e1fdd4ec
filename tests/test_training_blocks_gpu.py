"""GPU parity tests of the training building blocks against torch autograd / torch.optim on the CPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_loss_backward_matches_autograd():
    from hupr_b200 import ops
    from oracle import loss as ol
    torch.manual_seed(11)
    b = 3
    z1 = (torch.randn(b, 14, 64, 64) * 2).requires_grad_()
    z2 = (torch.randn(b, 14, 64, 64) * 2).requires_grad_()
    gt = torch.randint(0, 256, (b, 14, 2))
    gt[0, 0] = torch.tensor([300, 10])       # off-map joint: all-zero target
    targets = torch.from_numpy(np.stack([ol.generate_target(gt[i].numpy())[0] for i in range(b)]))
    p1, p2 = torch.sigmoid(z1), torch.sigmoid(z2)
    loss = F.binary_cross_entropy(p1, targets) + F.binary_cross_entropy(p2, targets)
    loss.backward()
    d1 = torch.full((b, 4096, 64), 7.0, device="cuda")
    d2 = torch.empty(b, 14, 64, 64, device="cuda")
    ops.heatmap_loss_bwd(p1.detach().cuda(), p2.detach().cuda(), gt, d1, d2)
    torch.cuda.synchronize()
    got1 = d1[..., :14].reshape(b, 64, 64, 14).permute(0, 3, 1, 2).cpu()
    scale = float(z1.grad.abs().max())
    assert float((got1 - z1.grad).abs().max()) < 2e-6 * scale and float((d2.cpu() - z2.grad).abs().max()) < 2e-6 * scale
    assert bool((d1[..., 14:] == 7.0).all())        # padding channels untouched


def test_adam_matches_torch_optim():
    from hupr_b200 import ops
    torch.manual_seed(12)
    n = 100003
    p0 = torch.randn(n)
    ref = p0.clone().requires_grad_()
    opt = torch.optim.Adam([ref], lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    p = p0.clone().cuda()
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    for step in range(1, 6):
        g = torch.randn(n) * (0.1 if step % 2 else 10.0)
        ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g.cuda(), m, v, step)
    torch.cuda.synchronize()
    err = float((p.cpu() - ref.detach()).abs().max())
    assert err < 1e-6                                     # parameters are O(1): this is ~2 ulp
    assert err < 5e-3 * float((ref.detach() - p0).abs().max())      # and well below the 5-step update itself (~5e-4)


@pytest.mark.parametrize("shape", [
    (2, 64, 128, 4, 32, 32, (3, 3, 3), (1, 1, 1)),      # 3-D conv
    (2, 64, 128, 8, 64, 64, (3, 3, 3), (1, 1, 1)),      # 3-D conv, halo-reuse kernel
    (1, 128, 64, 1, 32, 32, (1, 3, 3), (0, 1, 1)),      # decoder 2-D conv
    (2, 64, 64, 8, 64, 64, (8, 1, 1), (0, 0, 0)),       # temporal merge (depth-valid): dgrad pads depth by T-1
])
def test_conv_dgrad_via_flipped_filters_matches_autograd(shape):
    from hupr_b200.models.layers import pack_dgrad
    from hupr_b200.ops import SplitTensor, conv_gemm
    n, cin, cout, d, h, w, kernel, pad = shape
    torch.manual_seed(13)
    x = torch.randn(n, cin, d, h, w, device="cuda", dtype=torch.float64, requires_grad=True)
    wt = (torch.randn(cout, cin, *kernel, device="cuda", dtype=torch.float64) / (cin * kernel[0] * kernel[1] * kernel[2]) ** 0.5)
    y = F.conv3d(x, wt, padding=pad)
    dy = torch.randn_like(y)
    y.backward(dy)
    DY = SplitTensor.from_float(dy.float().permute(0, 2, 3, 4, 1).contiguous())
    WD = pack_dgrad(wt.float(), cout, cin)
    out = SplitTensor.empty((n, d, h, w, cin), "cuda")
    dpad = (kernel[0] - 1 - pad[0], pad[1], pad[2])
    conv_gemm(DY, cout, WD, cin, kernel=kernel, pad=dpad, out=out)
    torch.cuda.synchronize()
    got = out.float().permute(0, 4, 1, 2, 3).double()
    assert float((got - x.grad).abs().max() / x.grad.abs().max()) < 2e-5


@pytest.mark.parametrize("shape", [
    (2, 64, 128, 4, 32, 32, (3, 3, 3), (1, 1, 1)),
    (1, 64, 64, 2, 16, 16, (3, 3, 3), (1, 1, 1)),
    (2, 128, 64, 1, 32, 32, (1, 3, 3), (0, 1, 1)),
    (2, 64, 64, 1, 64, 64, (1, 1, 1), (0, 0, 0)),
    (2, 64, 64, 8, 64, 64, (8, 1, 1), (0, 0, 0)),       # temporal merge: depth-valid, dY lives at d = 0 of the input index space
])
def test_conv_wgrad_from_position_major_operands_matches_autograd(shape):
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    n, cin, cout, d, h, w, kernel, pad = shape
    torch.manual_seed(14)
    x = torch.randn(n, cin, d, h, w, device="cuda", dtype=torch.float64)
    wt = torch.randn(cout, cin, *kernel, device="cuda", dtype=torch.float64, requires_grad=True)
    y = F.conv3d(x, wt, padding=pad)
    dy = torch.randn_like(y)
    y.backward(dy)
    X = SplitTensor.from_float(x.float().permute(0, 2, 3, 4, 1).contiguous())
    DY = SplitTensor.from_float(dy.float().permute(0, 2, 3, 4, 1).contiguous())
    geom = ops.KMajorGeometry(n, d, h, w, pad)
    rows = -(-cout // 128) * 128
    xts = []
    for kw in range(kernel[2]):
        xt = SplitTensor.empty((cin, geom.ppad), "cuda", zero=True)
        ops.to_kmajor(X, 0, cin, geom, xt, shift=kw - pad[2])
        xts.append(xt)
    dyt = SplitTensor.empty((rows, geom.ppad), "cuda", zero=True)
    ops.to_kmajor(DY, 0, cout, geom, dyt)
    taps = kernel[0] * kernel[1] * kernel[2]
    out = torch.zeros(taps, rows, cin, device="cuda")
    ops.conv_wgrad(xts, cin, dyt, rows, geom, kernel, out)
    torch.cuda.synchronize()
    got = out[:, :cout].permute(1, 2, 0).reshape(cout, cin, *kernel).double()
    assert float((got - wt.grad).abs().max() / wt.grad.abs().max()) < 3e-5
