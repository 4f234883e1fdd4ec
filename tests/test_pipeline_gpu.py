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
    # the property is about batching, so both sides use the same arithmetic everywhere: the opt-in two-unit convolutions only engage where a
    # launch has enough tiles, i.e. on more layers at batch 32 than at batch 1
    net.quant_cross_terms = False
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


def test_batch_32_stream_matches_oracle_forward_on_all_32_windows():
    """The benched configuration (BASELINE.json configs[2], batch 32) against the oracle on ALL 32 windows: the torch-CPU oracle network is
    run on the stream's own standardised inputs (eight windows at a time); heat maps within 1e-3 element-wise relative (north_star), argmax
    keypoints bit-exact.  The per-frame-feature variant that bench.py times is tied to the same oracle through the window variant (2e-4)."""
    from hupr_b200.models import HuPRNet
    from hupr_b200.pipeline import RadarPoseStream
    from oracle import loss as ol
    from oracle import model as om
    from tests.test_model_gpu import make_cfg
    sd = om.make_state_dict(3)
    net = HuPRNet(make_cfg())
    net.load_state_dict(sd)
    net = net.cuda().eval()
    win = RadarPoseStream(net, 32, "cuda", use_graph=False, per_frame=False).prepare()
    gen = torch.Generator(device="cuda").manual_seed(5)
    win.adc.copy_(torch.randint(-2048, 2048, win.adc.shape, generator=gen, dtype=torch.int16, device="cuda"))
    kp = win.step().clone()
    heat, gcn = win.heatmap.reshape(32, 14, 64, 64).cpu(), win.gcn_heatmap.reshape(32, 14, 64, 64).cpu()
    vh, vv = win.vrdae_hori.cpu(), win.vrdae_vert.cpu()
    # the same step with the opt-in two-unit convolutions (HuPRNet.quant_cross_terms: fp16 main product + e4m3 cross terms)
    net.quant_cross_terms = True
    kp_q = win.step().clone()
    heat_q, gcn_q = win.heatmap.reshape(32, 14, 64, 64).cpu(), win.gcn_heatmap.reshape(32, 14, 64, 64).cpu()
    net.quant_cross_terms = False
    torch.cuda.synchronize()
    worst, worst_abs, smallest, worst_q = 0.0, 0.0, 1.0, 0.0
    for c in range(0, 32, 8):
        with torch.no_grad():
            ref_heat, ref_gcn = om.huprnet_forward(sd, vh[c:c + 8], vv[c:c + 8])
        ref_heat, ref_gcn = ref_heat.reshape(8, 14, 64, 64), ref_gcn.reshape(8, 14, 64, 64)
        e1 = float(((heat[c:c + 8] - ref_heat).abs() / ref_heat.abs()).max())
        e2 = float(((gcn[c:c + 8] - ref_gcn).abs() / ref_gcn.abs()).max())
        worst = max(worst, e1, e2)
        worst_abs = max(worst_abs, float((heat[c:c + 8] - ref_heat).abs().max()), float((gcn[c:c + 8] - ref_gcn).abs().max()))
        smallest = min(smallest, float(ref_heat.min()), float(ref_gcn.min()))
        assert e1 < 1e-3 and e2 < 1e-3, (c, e1, e2)
        ref_kp, _ = ol.get_max_preds(ref_gcn.numpy())
        assert np.array_equal(kp[c:c + 8].cpu().numpy(), ref_kp), c
        q1 = float(((heat_q[c:c + 8] - ref_heat).abs() / ref_heat.abs()).max())
        q2 = float(((gcn_q[c:c + 8] - ref_gcn).abs() / ref_gcn.abs()).max())
        worst_q = max(worst_q, q1, q2)
        assert q1 < 1e-3 and q2 < 1e-3, (c, q1, q2)
        assert np.array_equal(kp_q[c:c + 8].cpu().numpy(), ref_kp), c
    print("batch-32 stream vs oracle, all 32 windows: max element-wise relative heat-map error %.3g (max absolute error %.3g on maps in (0, 1); "
          "smallest reference value %.3g — the relative figure is set by the near-zero cells); with the two-unit convolutions %.3g"
          % (worst, worst_abs, smallest, worst_q))
    # the variant bench.py times (per-frame features, overlapping window views, CUDA graph) on the same ADC words
    pf = RadarPoseStream(net, 32, "cuda", use_graph=True, per_frame=True).prepare()
    pf.adc.copy_(win.adc)
    kp_pf = pf.step().clone()
    torch.cuda.synchronize()
    assert torch.equal(kp_pf, kp)
    assert float((pf.gcn_heatmap.reshape(32, 14, 64, 64).cpu() - gcn).abs().max()) < 2e-4 * float(gcn.abs().max())


def test_whole_chain_from_adc_words_matches_the_oracle_chain():
    """cascade -> window/Normalize -> network as ONE piece: int16 DCA1000 words through the GPU stream against the oracle chain
    (complex128 cascade restatement -> fp64 Normalize -> torch-CPU network) on the same frames.  The only substitution is the chirp slot
    fed by the clutter-removed Doppler-0 row (cube row 8 -> slot 4): it is round-off noise stretched to unit variance in BOTH
    implementations (SURVEY.md §7 trap 1), i.e. two unrelated random fields, so the oracle network is given the GPU's slot-4 planes; every
    other plane of its input comes from the oracle's own cascade + Normalize."""
    from oracle import cascade, loader
    from oracle import loss as ol
    from oracle import model as om
    n_windows = 2
    s, sd, hori, vert = build(n_windows, per_frame=False, use_graph=False)
    kp = s(hori.cuda(), vert.cuda()).clone()
    heat, gcn = s.heatmap.reshape(n_windows, 14, 64, 64).cpu(), s.gcn_heatmap.reshape(n_windows, 14, 64, 64).cpu()
    gpu_in = {0: s.vrdae_hori.cpu(), 1: s.vrdae_vert.cpu()}
    torch.cuda.synchronize()
    ref_in = {}
    for sensor in (0, 1):
        cubes = [cascade.generate_heatmap(cascade.synth_frame(f, sensor)) for f in range(s.n_frames)]
        x = np.stack([loader.vrdae_from_cubes(cubes[b:b + 8]) for b in range(n_windows)])
        signal = [c for c in range(8) if c != 4]
        # the oracle's own planes agree with the GPU's on every signal slot (cascade in complex64 + fp32 statistics vs complex128 + fp64)
        assert float(np.abs(x[:, :, signal] - gpu_in[sensor].numpy()[:, :, signal]).max()) < 2e-3
        x[:, :, 4] = gpu_in[sensor].numpy()[:, :, 4]
        ref_in[sensor] = torch.from_numpy(x)
    with torch.no_grad():
        ref_heat, ref_gcn = om.huprnet_forward(sd, ref_in[0], ref_in[1])
    ref_heat, ref_gcn = ref_heat.reshape(n_windows, 14, 64, 64), ref_gcn.reshape(n_windows, 14, 64, 64)
    e1 = float(((heat - ref_heat).abs() / ref_heat.abs()).max())
    e2 = float(((gcn - ref_gcn).abs() / ref_gcn.abs()).max())
    print("whole chain vs oracle chain: element-wise relative heat-map error %.3g / %.3g" % (e1, e2))
    assert e1 < 1e-3 and e2 < 1e-3
    ref_kp, _ = ol.get_max_preds(ref_gcn.numpy())
    assert np.array_equal(kp.cpu().numpy(), ref_kp)


def test_keypoint_sensitivity_to_the_doppler0_noise_plane():
    """The one AP-parity risk that can be quantified offline (SURVEY.md §7 trap 1 c): the network reads the clutter-removed Doppler-0 row
    (chirp slot 4) — pure round-off that Normalize stretches to unit variance — so its content differs between ANY two implementations
    (fp64 numpy vs complex64 kernels, even two BLAS builds).  Re-draw that slot as fresh unit-variance noise N times and measure how
    many of the 14 x B keypoints move and by how much the heat maps change.  With the seeded random weights of the tests this is a
    measurement, recorded in DESIGN.md §4; the assertions only bound what any usable network must satisfy (finite maps, keypoints that
    move stay within the 64 x 64 grid)."""
    s, sd, hori, vert = build(4, per_frame=False, use_graph=False)
    base_kp = s(hori.cuda(), vert.cuda()).clone()
    base_gcn = s.gcn_heatmap.clone()
    net = s.model
    vh, vv = s.vrdae_hori.clone(), s.vrdae_vert.clone()
    gen = torch.Generator(device="cuda").manual_seed(11)
    from hupr_b200 import ops
    moved, dist, rel = [], [], []
    for _ in range(8):
        h, v = vh.clone(), vv.clone()
        h[:, :, 4] = torch.randn(h[:, :, 4].shape, generator=gen, device="cuda")
        v[:, :, 4] = torch.randn(v[:, :, 4].shape, generator=gen, device="cuda")
        _, gcn = net(h, v)
        kp = ops.keypoints_argmax(gcn.reshape(4, 14, 64, 64))
        torch.cuda.synchronize()
        assert bool(torch.isfinite(gcn).all())
        assert float(kp.min()) >= 0 and float(kp.max()) <= 63
        d = (kp - base_kp).norm(dim=-1)
        moved.append(float((d > 0).float().mean()))
        dist.append(float(d.max()))
        rel.append(float((gcn.reshape(-1) - base_gcn.reshape(-1)).abs().max() / base_gcn.abs().max()))
    print("Doppler-0 noise plane re-drawn 8x (random seeded weights): keypoints moved %.1f %% on average (max %.1f %%), largest move %.1f heat-map "
          "pixels, heat-map change up to %.3g of the peak" % (100 * np.mean(moved), 100 * max(moved), max(dist), max(rel)))
