"""GPU parity tests of the model path (MNet, resampling, attention pieces, PRGCN, loss, whole HuPRNet forward) against the
torch-CPU oracle (oracle/model.py, oracle/loss.py, oracle/loader.py — each pinned to the reference by tests/golden/*.npz).

Tolerances (north_star: fp32 heatmaps within 1e-3 relative, integer index maps bit-exact):
  * per-op checks: <= 2e-5 of the tensor's max-abs (fp32-equivalent split-bf16 arithmetic);
  * whole-network heatmaps: max |got - ref| <= 1e-3 * max|ref| (measured margins are printed);
  * argmax keypoints: bit-exact."""
import os
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def make_cfg():
    ns = types.SimpleNamespace
    return ns(DATASET=ns(numFrames=8, rangeSize=64, heatmapSize=64, azimuthSize=64, elevationSize=8, numGroupFrames=8,
                         numKeypoints=14, imgSize=256),
              MODEL=ns(numFilters=32), TRAINING=ns(lossDecay=-1))


def rel_err(got, ref):
    got, ref = torch.as_tensor(got).double().cpu(), torch.as_tensor(ref).double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def cl_to_ncdhw(t, c):
    """SplitTensor [N, D, H, W, C'] -> float32 [N, c, D, H, W] on the CPU."""
    return t.float()[..., :c].permute(0, 4, 1, 2, 3).contiguous().cpu()


@pytest.fixture(scope="module")
def net_and_oracle():
    from hupr_b200.models import HuPRNet
    from oracle import model as om
    sd = om.make_state_dict(0)
    net = HuPRNet(make_cfg())
    net.load_state_dict(sd)
    net = net.cuda().eval()
    return net, sd


def test_window_normalize_matches_loader_oracle(golden_dir):
    from hupr_b200 import ops
    from oracle import cascade, loader
    cubes = [cascade.generate_heatmap(cascade.synth_frame(2, 0)), cascade.generate_heatmap(cascade.synth_frame(5, 1))]
    dev_cubes = torch.from_numpy(np.stack(cubes).astype(np.complex64)).cuda()
    slots = torch.tensor([1, 0, 0, 1, 1], dtype=torch.int32, device="cuda")
    got = ops.window_normalize(dev_cubes, slots)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    assert got.shape == (5, 8, 2, 64, 64, 8) and np.isfinite(got).all()
    ref0 = loader.vrdae_from_cubes([cubes[0]] * 8)[0]
    ref1 = loader.vrdae_from_cubes([cubes[1]] * 8)[0]
    sig = [c for c in range(8) if c != 4]        # slot 4 = Doppler-0 plane: round-off noise in both implementations
    for s, ref in ((0, ref1), (1, ref0), (2, ref0), (3, ref1), (4, ref1)):
        assert np.abs(got[s][sig] - ref[sig]).max() < 2e-4      # standardised planes are O(1); peaks reach ~40
        assert np.abs(got[s][sig] - ref[sig]).max() / np.abs(ref[sig]).max() < 1e-5
    # the noise plane must still be finite and standardised (mean 0, unbiased std 1), never NaN (SURVEY.md §7 trap 1)
    noise = got[:, 4].astype(np.float64)
    assert abs(noise.mean(axis=(2, 3))).max() < 1e-3 and abs(noise.std(axis=(2, 3), ddof=1) - 1).max() < 1e-3
    # golden fixture minted from the reference's own Normalize
    g = np.load(os.path.join(golden_dir, "loader_reference.npz"))
    assert np.abs(got[1][sig][:, :, ::4, ::4, :] - g["sample"][sig]).max() < 2e-4


def test_mnet_matches_oracle(net_and_oracle):
    from hupr_b200.ops import SplitTensor
    from oracle import model as om
    net, sd = net_and_oracle
    hori, vert = om.make_vrdae(2, 3)
    ra, re = net.forward_chirp(hori.cuda(), vert.cuda())
    torch.cuda.synchronize()
    ref_ra = om.chirp_net(hori, sd, "RAchirpNet")
    ref_re = om.chirp_net(vert, sd, "REchirpNet")
    assert rel_err(cl_to_ncdhw(ra, 32), ref_ra) < 2e-5
    assert rel_err(cl_to_ncdhw(re, 32), ref_re) < 2e-5


@pytest.mark.parametrize("shape,out", [((2, 64, 8, 64, 64), (4, 32, 32)), ((1, 128, 4, 32, 32), (2, 16, 16)),
                                       ((2, 128, 1, 16, 16), (1, 32, 32)), ((1, 64, 1, 32, 32), (1, 64, 64))])
def test_resample_matches_torch(shape, out):
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(4)
    x = torch.randn(shape)
    n, c, d, h, w = shape
    A = SplitTensor.from_float(x.permute(0, 2, 3, 4, 1).contiguous().cuda())
    O = SplitTensor.empty((n,) + out + (c + 64,), "cuda", zero=True)
    ops.resample_linear(A, c, O, out_ch_off=64)
    torch.cuda.synchronize()
    if d == 1:
        ref = F.interpolate(x[:, :, 0], size=out[1:], mode="bilinear", align_corners=True).unsqueeze(2)
    else:
        ref = F.interpolate(x, size=out, mode="trilinear", align_corners=True)
    got = O.float()
    assert rel_err(got[..., 64:].permute(0, 4, 1, 2, 3), ref) < 2e-5
    assert not got[..., :64].any()


def test_softmax_and_transpose():
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(5)
    for cols in (256, 1024, 4096):
        x = (torch.randn(37, cols) * 30).cuda()
        P = SplitTensor.empty((37, cols), "cuda")
        ops.softmax_rows(x, P)
        torch.cuda.synchronize()
        ref = F.softmax(x.double(), dim=1)
        assert float((P.float().double() - ref).abs().max()) < 1e-5      # hi+lo bf16 carries ~16 mantissa bits
    x = torch.randn(3, 1, 1, 256, 192).cuda()
    X = SplitTensor.from_float(x)
    T = SplitTensor.empty((3, 64, 256), "cuda")
    ops.transpose_split(X, 64, T, in_ch_off=128)
    torch.cuda.synchronize()
    assert torch.equal(T.hi, X.hi[:, 0, 0, :, 128:].transpose(1, 2)) and torch.equal(T.lo, X.lo[:, 0, 0, :, 128:].transpose(1, 2))


def test_prgcn_matches_oracle(net_and_oracle):
    from hupr_b200 import ops
    from oracle import model as om
    net, sd = net_and_oracle
    torch.manual_seed(6)
    for batch in (1, 5):
        logits = torch.randn(batch, 14, 64, 64) * 1.5
        ld = 64
        cl = torch.zeros(batch, 4096, ld)
        cl[..., :14] = logits.permute(0, 2, 3, 1).reshape(batch, 4096, 14)
        cl[..., 14:] = 123.0      # padding channels must be ignored
        pk = net._packed or net._pack()
        dw = pk["decoder"]
        ws = torch.empty(ops.prgcn_workspace_bytes(batch) // 4, device="cuda")
        heat = torch.empty(batch, 14, 64, 64, device="cuda")
        gcn = torch.empty(batch, 14, 64, 64, device="cuda")
        ops.prgcn_fwd(cl.cuda(), dw.gcn_w, dw.gcn_b, pk["adj"], ws, heat, gcn)
        torch.cuda.synchronize()
        ref_gcn = om.prgcn(logits, sd)[:, 0]
        assert float((heat.cpu() - torch.sigmoid(logits)).abs().max()) < 1e-6
        assert float((gcn.cpu() - ref_gcn).abs().max()) < 2e-5


def test_argmax_and_loss_match_oracle(golden_dir):
    from hupr_b200 import ops
    from oracle import loss as ol
    g = np.load(os.path.join(golden_dir, "loss_reference.npz"))
    gen = torch.Generator().manual_seed(int(g["seed"]))
    b = g["gt"].shape[0]
    heat = torch.rand((b, 14, 1, 64, 64), generator=gen) * 0.98 + 0.01
    gcn = torch.rand((b, 1, 14, 64, 64), generator=gen) * 0.98 + 0.01
    gt = torch.from_numpy(g["gt"])
    losses, gt2d, targets = ops.heatmap_loss_fwd(heat.view(b, 14, 64, 64).cuda(), gcn.view(b, 14, 64, 64).cuda(), gt, want_targets=True)
    pred2d = ops.keypoints_argmax(gcn.view(b, 14, 64, 64).cuda())
    torch.cuda.synchronize()
    total, loss2, ref_pred, ref_gt2d, ref_targets = ol.compute_loss(heat, gcn, g["gt"])
    losses = losses.cpu().numpy()
    assert abs(losses[0] - total) < 2e-6 * total and abs(losses[1] - loss2) < 2e-6 * loss2
    assert abs(losses[0] - float(g["loss"])) < 2e-6 * total          # the reference's own LossComputer output
    assert np.array_equal(pred2d.cpu().numpy(), ref_pred) and np.array_equal(pred2d.cpu().numpy(), g["pred2d"])     # bit-exact ints
    assert np.array_equal(gt2d.cpu().numpy(), ref_gt2d) and np.array_equal(gt2d.cpu().numpy(), g["gt2d"])
    assert np.abs(targets.cpu().numpy() - ref_targets).max() < 1e-6
    assert np.array_equal(targets.cpu().numpy() > 0, ref_targets > 0)
    # argmax ties: the first maximum wins (numpy.argmax)
    m = torch.zeros(2, 64, 64)
    m[0, 5, 7] = m[0, 9, 1] = 3.0
    m[1] = -1.0
    p = ops.keypoints_argmax(m.cuda()).cpu()
    assert p.tolist() == [[7.0, 5.0], [0.0, 0.0]]


@pytest.mark.parametrize("batch,seed,quant", [(1, 0, True), (2, 5, True), (2, 5, False)])
def test_huprnet_forward_matches_oracle_and_reference_golden(batch, seed, quant, golden_dir):
    """``quant``: the default arithmetic (two-unit 3-tap convolutions where the shape allows: fp16 main product + e4m3 cross terms) vs the
    three bf16 products everywhere; the same 1e-3 bar for the heat maps, a wider bound on the intermediate encoder features."""
    from hupr_b200 import ops
    from hupr_b200.models import HuPRNet
    from oracle import model as om
    from oracle.make_golden import MODEL_SAMPLE
    sd = om.make_state_dict(seed)
    net = HuPRNet(make_cfg())
    net.load_state_dict(sd)
    net = net.cuda().eval()
    net.quant_cross_terms = quant
    hori, vert = om.make_vrdae(batch, seed)
    heat, gcn = net(hori.cuda(), vert.cuda())
    torch.cuda.synchronize()
    assert heat.shape == (batch, 14, 1, 64, 64) and gcn.shape == (batch, 1, 14, 64, 64)
    with torch.no_grad():
        ref_heat, ref_gcn, inter = om.huprnet_forward(sd, hori, vert, return_intermediates=True)
    plan = net._plans[batch]
    for name, got, ref in (("enc_ra.f1", plan["enc_ra"].f1, inter["feats_ra"][0]), ("enc_ra.f2", plan["enc_ra"].f2, inter["feats_ra"][1]),
                           ("enc_ra.f3", plan["enc_ra"].f3, inter["feats_ra"][2]), ("enc_re.f3", plan["enc_re"].f3, inter["feats_re"][2])):
        e = rel_err(cl_to_ncdhw(got, ref.shape[1])[:, :, 0], ref)
        print("%s rel err %.3g" % (name, e))
        assert e < (4e-4 if quant else 1e-4), name
    logits = plan["dec"].logits_out[..., :14].reshape(batch, 64, 64, 14).permute(0, 3, 1, 2).cpu()
    print("logits rel err %.3g" % rel_err(logits, inter["logits"]))
    e_heat, e_gcn = rel_err(heat, ref_heat), rel_err(gcn, ref_gcn)
    print("heatmap rel err %.3g, gcn heatmap rel err %.3g" % (e_heat, e_gcn))
    assert e_heat < 1e-3 and e_gcn < 1e-3
    # element-wise relative error as well (heatmaps are sigmoid outputs well away from 0)
    assert float(((heat.cpu() - ref_heat).abs() / ref_heat.abs()).max()) < 1e-3
    assert float(((gcn.cpu() - ref_gcn).abs() / ref_gcn.abs()).max()) < 1e-3
    # golden samples minted from the reference HuPRNet itself
    g = np.load(os.path.join(golden_dir, "model_reference.npz"))
    key = "b%d_s%d" % (batch, seed)
    assert np.abs(heat.cpu().numpy()[MODEL_SAMPLE] - g[key + "_heatmap"]).max() < 1e-3 * np.abs(g[key + "_heatmap"]).max()
    assert np.abs(gcn.cpu().numpy()[MODEL_SAMPLE] - g[key + "_gcn"]).max() < 1e-3 * np.abs(g[key + "_gcn"]).max()
    # integer keypoints: bit-exact against the oracle's argmax
    from oracle import loss as ol
    pred = ops.keypoints_argmax(gcn.view(batch, 14, 64, 64)).cpu().numpy()
    ref_pred, _ = ol.get_max_preds(ref_gcn.view(batch, 14, 64, 64).numpy())
    assert np.array_equal(pred, ref_pred)


def test_forward_requires_eval_and_cuda():
    from hupr_b200.models import HuPRNet
    net = HuPRNet(make_cfg())
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 8, 8, 2, 64, 64, 8), torch.zeros(1, 8, 8, 2, 64, 64, 8))


@pytest.mark.parametrize("batch,s,residual,c", [(1, 128, False, 64), (2, 512, True, 64), (1, 4096, True, 64),
                                                (1, 128, True, 128), (3, 1024, False, 128)])
def test_fused_attention_matches_torch(batch, s, residual, c):
    """hupr_attention_fwd vs softmax(QK^T)V in float64; logits are un-scaled and large (|logit| up to ~60) like the network's."""
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(7)
    scale = 1.5 if c == 64 else 1.25       # logits std ~18 in both cases
    proj_q = torch.randn(batch, 1, 1, s, 4 * c, device="cuda") * scale
    proj_k = torch.randn(batch, 1, 1, s, 4 * c, device="cuda") * scale
    v = torch.randn(batch, 1, 1, s, c, device="cuda")
    PQ, PK, V = SplitTensor.from_float(proj_q), SplitTensor.from_float(proj_k), SplitTensor.from_float(v)
    VT = SplitTensor.empty((batch, c, s), "cuda")
    ops.transpose_split(V, c, VT)
    out = SplitTensor.empty((batch, 1, 1, s, 3 * c), "cuda", zero=True)
    ops.attention_fwd(PQ, 3 * c, PK, 2 * c, VT, c, out, c, residual=V if residual else None)
    torch.cuda.synchronize()
    q = PQ.float()[:, 0, 0, :, 3 * c:].double()
    k = PK.float()[:, 0, 0, :, 2 * c:3 * c].double()
    vv = V.float()[:, 0, 0].double()
    ref = torch.softmax(q @ k.transpose(1, 2), dim=-1) @ vv
    if residual:
        ref = ref + vv
    got = out.float()[:, 0, 0]
    err = rel_err(got[..., c:2 * c], ref)
    print("fused attention rel err %.3g" % err)
    assert err < 1e-4      # logits have std ~18 here: their 2^-17-relative product error shows up as a relative error of P
    assert not got[..., :c].any() and not got[..., 2 * c:].any()


def test_loss_computer_mirror_matches_reference_golden(golden_dir):
    """LossComputer(cfg, device).computeLoss keeps the reference's call signature and 4-tuple (losses.py:23-45)."""
    from hupr_b200.misc import LossComputer, get_max_preds
    g = np.load(os.path.join(golden_dir, "loss_reference.npz"))
    gen = torch.Generator().manual_seed(int(g["seed"]))
    b = g["gt"].shape[0]
    heat = (torch.rand((b, 14, 1, 64, 64), generator=gen) * 0.98 + 0.01).cuda()
    gcn = (torch.rand((b, 1, 14, 64, 64), generator=gen) * 0.98 + 0.01).cuda()
    lc = LossComputer(make_cfg(), torch.device("cuda"))
    loss, loss2, pred2d, gt2d = lc.computeLoss((heat, gcn), torch.from_numpy(g["gt"]))
    assert abs(float(loss) - float(g["loss"])) < 2e-6 * float(g["loss"]) and abs(float(loss2) - float(g["loss2"])) < 2e-6 * float(g["loss2"])
    assert pred2d.dtype == np.float32 and np.array_equal(pred2d, g["pred2d"]) and np.array_equal(gt2d, g["gt2d"])
    preds, maxvals = get_max_preds(gcn.view(b, 14, 64, 64))
    assert np.array_equal(preds, g["pred2d"]) and maxvals.shape == (b, 14, 1)


def test_c_abi_edge_cases_zero_sizes_alignment_and_bad_shapes():
    """Empty inputs are no-ops, misaligned pointers and unsupported shapes are rejected with the documented codes (no launch)."""
    import ctypes
    from hupr_b200 import _C, ops
    from hupr_b200.ops import SplitTensor
    lib = _C.lib()
    s = _C.stream_ptr()
    buf = torch.zeros(1 << 16, dtype=torch.float32, device="cuda")
    p = buf.data_ptr()
    assert lib.hupr_fft_cascade_i16(p, p, 0, s) == 0
    assert lib.hupr_window_normalize(p, p, 0, p, s) == 0 and lib.hupr_mnet_fwd(p, p, p, p, p, 0, s) == 0
    assert lib.hupr_softmax_rows(p, p, p, 0, 256, s) == 0 and lib.hupr_keypoints_argmax(p, 0, p, None, s) == 0
    assert lib.hupr_prgcn_workspace_bytes(0) == 0 and lib.hupr_frame_features_workspace_bytes(3) == 3 * 256 * 4
    assert lib.hupr_fft_cascade_i16(p + 2, p, 1, s) == -2                                   # HUPR_ERR_ALIGNMENT
    assert lib.hupr_softmax_rows(p, p, p, 4, 4100, s) == -1 and lib.hupr_softmax_rows(p, p, p, 4, 6, s) == -1      # cols > 4096 / not % 4
    assert lib.hupr_heatmap_loss_fwd(p, p, p, 2, p, 8, p, None, None, s) == -5             # HUPR_ERR_WORKSPACE
    assert lib.hupr_adam_step(p, p, p, p, 16, 1e-4, 0.9, 0.999, 1e-8, 1e-4, 0, None, None, s) == -1      # step < 1 without a device counter
    d = _C.AttnDesc()
    x = SplitTensor.empty((1, 1, 1, 128, 96), "cuda", zero=True)
    d.q_hi = d.k_hi = d.vt_hi = d.o_hi = x.hi.data_ptr()
    d.q_lo = d.k_lo = d.vt_lo = d.o_lo = x.lo.data_ptr()
    d.q_ld = d.k_ld = d.o_ld = 96
    d.batch, d.s, d.c = 1, 128, 96
    assert lib.hupr_attention_fwd(ctypes.byref(d), s) == -1                                 # head dim 96 is not supported
    with pytest.raises(RuntimeError, match="bad argument"):
        ops.resample_linear(x, 96, SplitTensor.empty((1, 1, 1, 64, 100), "cuda"))          # ld not a multiple of 8
    torch.cuda.synchronize()
