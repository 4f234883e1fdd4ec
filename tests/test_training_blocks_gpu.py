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
    xt = SplitTensor.empty((kernel[2] * cin, geom.ppad), "cuda", zero=True)
    for kw in range(kernel[2]):
        ops.to_kmajor(X, 0, cin, geom, SplitTensor(xt.hi[kw * cin:(kw + 1) * cin], xt.lo[kw * cin:(kw + 1) * cin]), shift=kw - pad[2])
    dyt = SplitTensor.empty((rows, geom.ppad), "cuda", zero=True)
    ops.to_kmajor(DY, 0, cout, geom, dyt)
    taps = kernel[0] * kernel[1] * kernel[2]
    out = torch.zeros(kernel[0] * kernel[1], rows, kernel[2] * cin, device="cuda")
    ops.conv_wgrad(xt, cin, dyt, rows, geom, kernel, out)
    torch.cuda.synchronize()
    out = out.view(kernel[0] * kernel[1], rows, kernel[2], cin).permute(0, 2, 1, 3).reshape(taps, rows, cin)
    got = out[:, :cout].permute(1, 2, 0).reshape(cout, cin, *kernel).double()
    assert float((got - wt.grad).abs().max() / wt.grad.abs().max()) < 3e-5


def _cl(x):
    """NCDHW float -> SplitTensor channels-last."""
    from hupr_b200.ops import SplitTensor
    return SplitTensor.from_float(x.float().permute(0, 2, 3, 4, 1).contiguous().cuda())


def _nc(t):
    return t.float().permute(0, 4, 1, 2, 3).cpu().double()


def test_batchnorm_train_forward_backward_match_autograd():
    """relu(BN(z)) with batch statistics, forward and backward, assembled from hupr_channel_sums / hupr_affine_act /
    hupr_bn_bwd_apply exactly as the training step does (layers.py:45-53 in training mode)."""
    from hupr_b200 import train_ops as T
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(15)
    n, c, d, h, w = 2, 64, 4, 16, 16
    z = (torch.randn(n, c, d, h, w, dtype=torch.float64) * 1.7 + 0.3).requires_grad_()
    gamma = (torch.rand(c, dtype=torch.float64) + 0.5).requires_grad_()
    beta = (torch.randn(c, dtype=torch.float64) * 0.2).requires_grad_()
    y = F.relu(F.batch_norm(z, None, None, gamma, beta, True, 0.1, 1e-5))
    gy = torch.randn_like(y)
    y.backward(gy)
    N = n * d * h * w
    Z = SplitTensor.empty((n, d, h, w, 2 * c), "cuda", zero=True)         # z lives in channels [c, 2c) of a wider buffer
    Zsrc = _cl(z.detach())
    Z.hi[..., c:] = Zsrc.hi; Z.lo[..., c:] = Zsrc.lo
    s1 = torch.zeros(c, dtype=torch.float64, device="cuda"); s2 = torch.zeros_like(s1)
    T.channel_sums(T.SUMS_STATS, (Z, c), c, s1, s2)
    mean = s1 / N
    var = s2 / N - mean * mean
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    scale = (gamma.detach().cuda() * rstd).float(); shift = (beta.detach().cuda() - mean * gamma.detach().cuda() * rstd).float()
    Y = SplitTensor.empty((n, d, h, w, c), "cuda")
    T.affine_act((Z, c), c, Y, scale1=scale, shift1=shift, slope=torch.zeros(c, device="cuda"))
    torch.cuda.synchronize()
    assert float((_nc(Y) - y.detach()).abs().max()) < 3e-5 * float(y.abs().max())
    G = _cl(gy)
    t1 = torch.zeros_like(s1); t2 = torch.zeros_like(s1)
    meanf, rstdf = mean.float(), rstd.float()
    T.channel_sums(T.SUMS_BN_BWD, G, c, t1, t2, b=(Z, c), mask=Y, mean=meanf, rstd=rstdf)
    DZ = SplitTensor.empty((n, d, h, w, c), "cuda")
    T.bn_bwd_apply(G, (Z, c), c, meanf, rstdf, scale, (t1 / N).float(), (t2 / N).float(), DZ, mask=Y)
    torch.cuda.synchronize()
    assert float((t1.cpu() - beta.grad).abs().max()) < 1e-4 * float(beta.grad.abs().max())
    assert float((t2.cpu() - gamma.grad).abs().max()) < 1e-4 * float(gamma.grad.abs().max())
    assert float((_nc(DZ) - z.grad).abs().max()) < 5e-5 * float(z.grad.abs().max())


def test_prelu_resample_softmax_backward_match_autograd():
    from hupr_b200 import train_ops as T
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(16)
    # PReLU backward + slope gradient
    n, c, d, h, w = 2, 64, 1, 16, 16
    s = torch.randn(n, c, d, h, w, dtype=torch.float64, requires_grad=True)
    a = torch.tensor([0.3], dtype=torch.float64, requires_grad=True)
    y = F.prelu(s, a); g = torch.randn_like(y); y.backward(g)
    S, G = _cl(s.detach()), _cl(g)
    DS = SplitTensor.empty((n, d, h, w, c), "cuda")
    T.act_bwd(G, S, c, torch.full((c,), 0.3, device="cuda"), DS)
    p1 = torch.zeros(c, dtype=torch.float64, device="cuda"); p2 = torch.zeros_like(p1)
    T.channel_sums(T.SUMS_PRELU, G, c, p1, p2, b=S)
    torch.cuda.synchronize()
    assert float((_nc(DS) - s.grad).abs().max()) < 3e-5 * float(s.grad.abs().max())
    assert abs(float(p1.sum()) - float(a.grad)) < 1e-4 * abs(float(a.grad)) and float((p2.cpu() - g.sum(dim=(0, 2, 3, 4))).abs().max()) < 1e-3
    # resample backward (trilinear 0.5x and bilinear 2x, align_corners)
    for shape, out in (((2, 64, 8, 32, 32), (4, 16, 16)), ((1, 64, 1, 16, 16), (1, 32, 32))):
        x = torch.randn(shape, dtype=torch.float64, requires_grad=True)
        if shape[2] == 1:
            yy = F.interpolate(x[:, :, 0], size=out[1:], mode="bilinear", align_corners=True).unsqueeze(2)
        else:
            yy = F.interpolate(x, size=out, mode="trilinear", align_corners=True)
        gg = torch.randn_like(yy); yy.backward(gg)
        din = torch.zeros((shape[0], shape[2], shape[3], shape[4], shape[1]), device="cuda")
        T.resample_linear_bwd(_cl(gg), shape[1], din)
        torch.cuda.synchronize()
        assert float((din.permute(0, 4, 1, 2, 3).cpu().double() - x.grad).abs().max()) < 3e-5 * float(x.grad.abs().max())
    # softmax backward
    logits = (torch.randn(40, 1024, dtype=torch.float64) * 3).requires_grad_()
    p = torch.softmax(logits, dim=1); gp = torch.randn_like(p); p.backward(gp)
    P = SplitTensor.from_float(p.detach().float().cuda())
    DSm = SplitTensor.empty((40, 1024), "cuda")
    T.softmax_bwd_rows(P, gp.float().cuda(), DSm)
    torch.cuda.synchronize()
    assert float((DSm.float().cpu().double() - logits.grad).abs().max()) < 5e-5 * float(logits.grad.abs().max())
    # accumulate: split + split + float -> split
    A, B = SplitTensor.from_float(torch.randn(4, 8, 64).cuda()), SplitTensor.from_float(torch.randn(4, 8, 64).cuda())
    f = torch.randn(4, 8, 128, device="cuda")
    O = SplitTensor.empty((4, 8, 64), "cuda")
    T.accumulate(O, 64, a=A, b=B, f=f, f_off=64)
    torch.cuda.synchronize()
    assert float((O.float() - (A.float() + B.float() + f[..., 64:])).abs().max()) < 1e-4


def test_attention_backward_fused_epilogues_match_float64():
    """The pieces that rebuild P and dS inside GEMM epilogues (layers.py:126-133 under autograd): the fused forward's saved
    log-sum-exp, P = exp(QK^T - lse) (hupr_conv_desc.row_mode 1), rowdot = <dO, O - residual> (hupr_rowdot) and
    dS = P * (dO V^T - rowdot) (row_mode 2), against float64 softmax algebra."""
    from hupr_b200 import ops
    from hupr_b200 import train_ops as T
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(31)
    b, s, c = 2, 256, 64
    q = torch.randn(b, s, c, device="cuda") * 0.6
    k = torch.randn(b, s, c, device="cuda") * 0.6
    v = torch.randn(b, s, c, device="cuda")
    do = torch.randn(b, s, c, device="cuda")
    Q, K, V, DO = (SplitTensor.from_float(t.view(b, 1, 1, s, c)) for t in (q, k, v, do))
    qd, kd, vd, dod = (SplitTensor.from_float(t).float().double() for t in (q, k, v, do))
    logits = qd @ kd.transpose(1, 2)
    p_ref = torch.softmax(logits, dim=2)
    o_ref = p_ref @ vd + vd                                          # cross branch: residual V
    # forward: output and log-sum-exp
    vt = ops.transpose_split(V, c, SplitTensor.empty((b, c, s), "cuda"))
    out = SplitTensor.empty((b, 1, 1, s, c), "cuda")
    lse = torch.empty((b, s), device="cuda")
    ops.attention_fwd(Q, 0, K, 0, vt, c, out, 0, residual=V, lse=lse)
    torch.cuda.synchronize()
    assert float((lse.double() - torch.logsumexp(logits, dim=2)).abs().max()) < 6e-5      # the logits themselves carry ~2e-5 (|q||k| 2^-17)
    assert float((out.float().double().view(b, s, c) - o_ref).abs().max()) < 1e-4
    # P = exp(logits - lse)
    probs = SplitTensor.empty((b, 1, 1, s, s), "cuda")
    ops.conv_gemm(Q, c, SplitTensor(K.hi.view(b, s, c), K.lo.view(b, s, c)), s, w_batched=True, out=probs, row_vec=lse, row_mode=1)
    torch.cuda.synchronize()
    assert float((probs.float().double().view(b, s, s) - p_ref).abs().max()) < 2e-5 * float(p_ref.max())
    # rowdot and dS
    rd = T.rowdot(DO, out, c, torch.empty((b, s), device="cuda"), sub=V)
    dp_ref = dod @ vd.transpose(1, 2)
    rd_ref = (p_ref * dp_ref).sum(dim=2)
    torch.cuda.synchronize()
    assert float((rd.double() - rd_ref).abs().max()) < 1e-4 * float(rd_ref.abs().max())
    ds = SplitTensor.empty((b, 1, 1, s, s), "cuda")
    ops.conv_gemm(DO, c, SplitTensor(V.hi.view(b, s, c), V.lo.view(b, s, c)), s, w_batched=True, out=ds, residual=probs, row_vec=rd, row_mode=2)
    torch.cuda.synchronize()
    ds_ref = p_ref * (dp_ref - rd_ref.unsqueeze(2))
    assert float((ds.float().double().view(b, s, s) - ds_ref).abs().max()) < 1e-4 * float(ds_ref.abs().max())
    # the row modes refuse epilogue combinations they do not define
    with pytest.raises(RuntimeError):
        ops.conv_gemm(DO, c, SplitTensor(V.hi.view(b, s, c), V.lo.view(b, s, c)), s, w_batched=True, out=ds, row_vec=rd, row_mode=2)
