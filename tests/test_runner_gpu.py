"""GPU test of the drop-in entry points (Runner / dataset / checkpoints) on a synthetic HuPR-layout directory."""
import json
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_runner_trains_evaluates_and_resumes(tmp_path, monkeypatch):
    from hupr_b200 import config
    from hupr_b200.tools import Runner
    from oracle import loader
    from tests.test_host_mirrors import make_dataset_dir
    root = str(tmp_path / "data")
    for phase, groups in (("train", [3]), ("val", [5])):
        blocks = make_dataset_dir(root, groups, 4, seed=len(phase))
        with open(os.path.join(root, "hrnet_annot_%s.json" % phase), "w") as fp:
            json.dump(blocks, fp)
    monkeypatch.chdir(tmp_path)
    os.makedirs("logs", exist_ok=True)
    cfg = config.default_config(DATASET__dataDir=root, DATASET__trainName=[3], DATASET__valName=[5], DATASET__duration=4,
                                TRAINING__batchSize=2, TRAINING__epochs=1, TEST__batchSize=4)
    args = types.SimpleNamespace(seed=0, dir="run", visDir="none", gpuIDs=[0], eval=False, sampling_ratio=1, keypoints=False)
    runner = Runner(args, cfg)
    # loader item: standardised planes equal the reference Normalize arithmetic (fp64 oracle) on the signal planes
    item = runner.trainSet[2]
    assert item["VRDAEmap_hori"].shape == (8, 8, 2, 64, 64, 8) and item["VRDAEmap_hori"].is_cuda and item["jointsGroup"].shape == (14, 2)
    cubes = [np.load(runner.trainSet.VRDAEPaths_hori[i]) for i in (0, 0, 0, 1, 2, 3, 3, 3)]      # window of index 2 in a 4-frame capture
    ref = loader.vrdae_from_cubes(cubes)
    assert np.abs(item["VRDAEmap_hori"].cpu().numpy() - ref).max() < 2e-4
    runner.loadModelWeight("checkpoint")            # nothing to load yet
    before = torch.cat([p.detach().reshape(-1) for p in runner.model.parameters()]).clone()
    runner.train()
    after = torch.cat([p.detach().reshape(-1) for p in runner.model.parameters()])
    assert float((after - before).abs().max()) > 0
    results = json.load(open("logs/run/val_results.json"))
    assert len(results) == 4 and len(results[0]["keypoints"]) == 42 and results[0]["image_id"] == 500000
    losses = json.load(open("logs/run/train_loss_list_0.json"))
    assert len(losses) == 2 and all(np.isfinite(losses))
    ckpt = torch.load("logs/run/checkpoint.pth", weights_only=False)
    assert sorted(ckpt) == ["accuracy", "epoch", "model_state_dict", "optimizer_state_dict"] and len(ckpt["model_state_dict"]) == 255
    assert sorted(ckpt["optimizer_state_dict"]) == ["param_groups", "state"] and len(ckpt["optimizer_state_dict"]["state"]) == 165
    # the optimizer state is torch.optim.Adam's own format: torch can load it
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros_like(p)) for p in runner.model.parameters()], lr=1e-4, weight_decay=1e-4)
    opt.load_state_dict(ckpt["optimizer_state_dict"])
    # resume: a fresh Runner restores weights, moments, step count and epoch
    resumed = Runner(args, cfg)
    resumed.loadModelWeight("checkpoint")
    assert resumed.start_epoch == 0 and resumed.trainer.step_count == runner.trainer.step_count == 2
    assert torch.equal(torch.cat([p.detach().reshape(-1) for p in resumed.model.parameters()]), after)
    assert torch.equal(resumed.trainer.exp_avg, runner.trainer.exp_avg)
    # evaluation entry point writes test_results.json from the val-style directory reused as the test split
    eval_cfg = config.default_config(DATASET__dataDir=root, DATASET__testName=[5], DATASET__duration=4, TEST__batchSize=4)
    os.replace(os.path.join(root, "hrnet_annot_val.json"), os.path.join(root, "hrnet_annot_test.json"))
    eval_args = types.SimpleNamespace(seed=0, dir="run", visDir="none", gpuIDs=[0], eval=True, sampling_ratio=1, keypoints=False)
    ev = Runner(eval_args, eval_cfg)
    ev.loadModelWeight("checkpoint")
    ap = ev.eval()
    assert len(json.load(open("logs/run/test_results.json"))) == 4
    # AP from the device-resident poses (hupr_keypoint_oks) == AP of the written results file through the host evaluator, and the
    # per-joint variant runs; a random-init network scores ~0 but the statistic is a finite number in [0, 1]
    assert 0.0 <= ap <= 1.0
    assert ev.testSet.evaluate("logs/run") == pytest.approx(ap, abs=1e-12)
    assert 0.0 <= ev.testSet.evaluateEach("logs/run") <= 1.0
