"""CPU tests: the model / loss / loader oracles against golden vectors minted from the reference's own modules
(oracle/make_golden.py model|loss|loader, run in the build container where /root/reference is mounted)."""
import os

import numpy as np
import torch

from oracle import cascade, loader, loss
from oracle import model as om
from oracle.make_golden import MODEL_SAMPLE, tensor_stats


def test_state_dict_layout():
    spec = om.state_dict_spec()
    assert len(spec) == 255
    assert sum(int(np.prod(s)) for n, s in spec if "running" not in n and "num_batches" not in n) == 35542668


def test_model_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_reference.npz"))
    batch, seed = 1, 0
    sd = om.make_state_dict(seed)
    hori, vert = om.make_vrdae(batch, seed)
    with torch.no_grad():
        heat, gcn, inter = om.huprnet_forward(sd, hori, vert, return_intermediates=True)
    key = "b%d_s%d" % (batch, seed)
    assert heat.shape == (batch, 14, 1, 64, 64) and gcn.shape == (batch, 1, 14, 64, 64)
    np.testing.assert_allclose(heat.numpy()[MODEL_SAMPLE], g[key + "_heatmap"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(gcn.numpy()[MODEL_SAMPLE], g[key + "_gcn"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(tensor_stats(heat), g[key + "_heatmap_stats"], rtol=1e-6)
    np.testing.assert_allclose(tensor_stats(gcn), g[key + "_gcn_stats"], rtol=1e-6)
    np.testing.assert_allclose(tensor_stats(inter["ra"]), g[key + "_chirp_stats"], rtol=1e-6)
    for i, f in enumerate(inter["feats_ra"]):
        np.testing.assert_allclose(f.numpy()[:, ::8, ::5, ::5], g[key + "_enc%d" % i], rtol=0, atol=1e-5)


def test_loss_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "loss_reference.npz"))
    gen = torch.Generator().manual_seed(int(g["seed"]))
    b = g["gt"].shape[0]
    heat = torch.rand((b, 14, 1, 64, 64), generator=gen) * 0.98 + 0.01
    gcn = torch.rand((b, 1, 14, 64, 64), generator=gen) * 0.98 + 0.01
    total, loss2, pred2d, gt2d, targets = loss.compute_loss(heat, gcn, g["gt"])
    assert abs(total - float(g["loss"])) < 1e-6 and abs(loss2 - float(g["loss2"])) < 1e-6
    assert np.array_equal(pred2d, g["pred2d"]) and np.array_equal(gt2d, g["gt2d"])      # integer keypoints: bit-exact
    assert np.array_equal(targets[:, :, ::3, ::3], g["targets_sample"])
    assert int((targets > 0).sum()) == int(g["targets_nonzero"])
    assert abs(targets.astype(np.float64).sum() - float(g["targets_sum"])) < 1e-9


def test_loader_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "loader_reference.npz"))
    frame_idx, sensor = (int(v) for v in g["frame"])
    cube = cascade.generate_heatmap(cascade.synth_frame(frame_idx, sensor))
    out = loader.vrdae_from_cubes([cube] * 8)
    assert np.array_equal(out[0], out[7])
    # Doppler row 8 (slot 4) is the clutter-removed DC bin: round-off noise stretched to unit variance (SURVEY.md §7 trap 1);
    # it is reproducible only because oracle and reference run the same fp64 operations in the same order
    np.testing.assert_allclose(out[0][:, :, ::4, ::4, :], g["sample"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(out[0].astype(np.float64).sum(axis=(2, 3)), g["sums"], atol=1e-3)
    np.testing.assert_allclose((out[0].astype(np.float64) ** 2).sum(axis=(2, 3)), g["sqsums"], rtol=1e-5)


def test_window_indices_equal_clamp_closed_form():
    """dataset.py:126-138 is clamp(index - 4 + j, first frame of the capture, last frame of the capture)."""
    for index in list(range(0, 1200)) + [165599 - k for k in range(10)]:
        first = index - index % loader.DURATION
        last = first + loader.DURATION - 1
        expect = [min(max(index - 4 + j, first), last) for j in range(8)]
        assert loader.window_indices(index) == expect


def test_training_oracle_matches_reference_train_golden(golden_dir):
    """oracle.model.training_gradients (train-mode forward + BCE + autograd) against ONE training pass of the unmodified reference
    (HuPRNet.train(), LossComputer.computeLoss, loss.backward() — tools/run.py:66,76-78), minted by `make_golden train`: losses and
    every one of the 165 parameter gradients (strided samples + checksums)."""
    from oracle.make_golden import GRAD_STRIDE, train_case_inputs
    g = np.load(os.path.join(golden_dir, "train_reference.npz"))
    assert int(g["grad_stride"]) == GRAD_STRIDE
    sd, hori, vert, joints = train_case_inputs()
    assert np.array_equal(joints.numpy(), g["joints"])
    total, loss2, grads = om.training_gradients(sd, hori, vert, joints.numpy())
    assert abs(total - float(g["loss"])) < 2e-6 and abs(loss2 - float(g["loss2"])) < 2e-6
    names = [str(n) for n in g["names"]]
    assert len(names) == 165 and sorted(names) == sorted(grads)
    for name in names:
        got = grads[name].reshape(-1)
        ref_stats = g["s/" + name]
        scale = max(float(ref_stats[2]), 1e-30)
        assert float(np.abs(got[::GRAD_STRIDE].numpy() - g["g/" + name]).max()) <= 1e-5 * scale + 1e-9, name
        np.testing.assert_allclose(tensor_stats(got)[1:], ref_stats[1:], rtol=1e-4, err_msg=name)
