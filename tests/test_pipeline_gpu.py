"""GPU tests of the end-to-end stream (DCA1000 words -> keypoints): the two pre-encoder variants agree, and the stream's network
stage matches the oracle model evaluated on the stream's own network inputs.

Note (SURVEY.md §7 trap 1): network chirp slot 4 is the clutter-removed DC Doppler bin — round-off noise that `Normalize`
stretches to unit variance.  fp32 (GPU) and fp64 (reference) noise are different random fields, so an end-to-end comparison must
feed the oracle the GPU's standardised planes; the signal planes are compared against the reference in test_model_gpu.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def build(n_windows, per_frame, use_graph):
    from hupr_b200.models import HuPRNet
    from hupr_b200.pipeline import RadarPoseStream
    from oracle import cascade
    from oracle import model as om
    from tests.test_model_gpu import make_cfg
    sd = om.make_state_dict(2)
    net = HuPRNet(make_cfg())
    net.load_state_dict(sd)
    net = net.cuda().eval()
    stream = RadarPoseStream(net, n_windows, "cuda", use_graph=use_graph, per_frame=per_frame).prepare()
    n = stream.n_frames
    hori = np.stack([cascade.complex_to_dca1000(cascade.synth_frame(f, 0)) for f in range(n)])
    vert = np.stack([cascade.complex_to_dca1000(cascade.synth_frame(f, 1)) for f in range(n)])
    return stream, sd, torch.from_numpy(hori), torch.from_numpy(vert)


def test_stream_variants_agree_and_match_oracle_model():
    from oracle import loss as ol
    from oracle import model as om
    s_win, sd, hori, vert = build(2, per_frame=False, use_graph=False)
    kp_win = s_win(hori.cuda(), vert.cuda()).clone()
    heat_win, gcn_win = s_win.heatmap.clone(), s_win.gcn_heatmap.clone()
    torch.cuda.synchronize()
    # oracle network on the stream's own standardised inputs
    with torch.no_grad():
        ref_heat, ref_gcn = om.huprnet_forward(sd, s_win.vrdae_hori.cpu(), s_win.vrdae_vert.cpu())
    assert float((heat_win.cpu() - ref_heat).abs().max() / ref_heat.abs().max()) < 1e-3
    assert float((gcn_win.cpu() - ref_gcn).abs().max() / ref_gcn.abs().max()) < 1e-3
    ref_kp, _ = ol.get_max_preds(ref_gcn.view(2, 14, 64, 64).numpy())
    assert np.array_equal(kp_win.cpu().numpy(), ref_kp)
    # per-frame features + overlapping window views (CUDA graph replay) give the same poses
    s_pf, _, _, _ = build(2, per_frame=True, use_graph=True)
    kp_pf = s_pf(hori.pin_memory(), vert.pin_memory()).clone()
    torch.cuda.synchronize()
    assert float((s_pf.heatmap - heat_win).abs().max()) < 1e-4 and float((s_pf.gcn_heatmap - gcn_win).abs().max()) < 1e-4
    assert torch.equal(kp_pf, kp_win)
    # replaying the graph on the same input is deterministic
    kp_again = s_pf.step().clone()
    torch.cuda.synchronize()
    assert torch.equal(kp_again, kp_pf)


def test_frame_features_equal_window_path_per_frame():
    """hupr_plane_stats + hupr_frame_features == hupr_window_normalize + hupr_mnet_fwd on the same cubes (1e-5)."""
    from hupr_b200 import ops
    from hupr_b200.ops import SplitTensor
    torch.manual_seed(3)
    cubes = torch.view_as_complex(torch.randn(3, 16, 64, 64, 8, 2, device="cuda") * 1e4 + 500.0)
    w = torch.randn(32, 2, 2, device="cuda") * 0.5
    b = torch.randn(32, device="cuda") * 0.1
    slots = torch.arange(3, dtype=torch.int32, device="cuda")
    vr = ops.window_normalize(cubes, slots)
    ref = SplitTensor.empty((3, 64, 64, 32), "cuda")
    ops.mnet_fwd(vr, w.contiguous(), b, ref)
    stats = ops.plane_stats(cubes)
    got = SplitTensor.empty((2, 64, 64, 32), "cuda")
    ops.frame_features(cubes, stats, 1, 2, w.contiguous(), b, got)
    torch.cuda.synchronize()
    assert float((got.float() - ref.float()[1:]).abs().max()) < 2e-5


def test_prefetched_ingest_equals_direct_call():
    s, _, hori, vert = build(2, per_frame=True, use_graph=True)
    direct = s(hori.pin_memory(), vert.pin_memory()).clone()
    torch.cuda.synchronize()
    s.adc.zero_()
    hp, vp = hori.pin_memory(), vert.pin_memory()
    s.prefetch(hp, vp)
    first = s.step_prefetched().clone()
    s.prefetch(vp, hp)                      # swapped sensors: a different result must come out of the second step
    second = s.step_prefetched().clone()
    torch.cuda.synchronize()
    assert torch.equal(first, direct)
    swapped = s(vert.pin_memory(), hori.pin_memory()).clone()
    torch.cuda.synchronize()
    assert torch.equal(second, swapped)


def test_batch_32_stream_equals_single_window_streams():
    """Full-size property (BASELINE.json configs[2] shape): the 32 poses of one batch-32 step equal the poses of 32 separate
    single-window steps over the same frames — windows are independent in eval mode, so batching may only change the summation order
    inside the small-grid split-K contractions, which the hi/lo rounding of ~40 layers amplifies to a few 1e-5 (the same size as the
    whole-network error against the oracle; bound 2e-4, five times inside the 1e-3 bar) — argmax keypoints identical."""
    from hupr_b200.models import HuPRNet
    from hupr_b200.pipeline import RadarPoseStream
    from oracle import model as om
    from tests.test_model_gpu import make_cfg
    net = HuPRNet(make_cfg())
    net.load_state_dict(om.make_state_dict(3))
    net = net.cuda().eval()
    big = RadarPoseStream(net, 32, "cuda", use_graph=False).prepare()
    gen = torch.Generator(device="cuda").manual_seed(5)
    big.adc.copy_(torch.randint(-2048, 2048, big.adc.shape, generator=gen, dtype=torch.int16, device="cuda"))
    kp_big = big.step().clone()
    heat_big = big.gcn_heatmap.reshape(32, 14, 64, 64).clone()
    one = RadarPoseStream(net, 1, "cuda", use_graph=False).prepare()
    nb, n1 = big.n_frames, one.n_frames
    for w in (0, 13, 31):
        one.adc[:n1].copy_(big.adc[w:w + n1])
        one.adc[n1:].copy_(big.adc[nb + w:nb + w + n1])
        kp = one.step()
        torch.cuda.synchronize()
        assert torch.equal(kp[0], kp_big[w]), w
        h1 = one.gcn_heatmap.reshape(14, 64, 64)
        assert float((h1 - heat_big[w]).abs().max()) < 2e-4 * float(heat_big[w].abs().max()), w
