"""CPU tests of the host-side mirrors that carry no arithmetic: configuration, CLI flags, GT-annotation synthesis, window index map,
dataset bookkeeping (no GPU: items are not materialised here)."""
import json
import os
import types

import numpy as np
import pytest

from oracle import loader


def make_dataset_dir(root, groups, frames, seed=0):
    """Synthetic HuPR-layout directory: random cubes + hrnet-style annotations for one phase ('train')."""
    rng = np.random.default_rng(seed)
    blocks = []
    for g in groups:
        for sensor in ("hori", "vert"):
            os.makedirs(os.path.join(root, "single_%d" % g, sensor), exist_ok=True)
        seq = []
        for f in range(frames):
            for sensor in ("hori", "vert"):
                cube = (rng.standard_normal((16, 64, 64, 8)) + 1j * rng.standard_normal((16, 64, 64, 8))) * 1e4
                np.save(os.path.join(root, "single_%d" % g, sensor, "%09d.npy" % f), cube.astype(np.complex64))
            joints = rng.integers(20, 236, (14, 2)).tolist()
            seq.append({"image": "%09d.jpg" % f, "joints": joints, "bbox": [10.0, 20.0, 200.0, 240.0]})
        blocks.append(seq)
    return blocks


def test_default_config_and_yaml_round_trip(tmp_path):
    from hupr_b200 import config
    cfg = config.default_config()
    assert cfg.DATASET.numKeypoints == 14 and cfg.MODEL.numFilters == 32 and cfg.TRAINING.lossDecay == -1
    assert len(cfg.DATASET.trainName) == 193 and len(cfg.DATASET.valName) == 21 and len(cfg.DATASET.testName) == 21
    assert cfg.DATASET.trainName[:3] == [2, 3, 4] and cfg.DATASET.trainName[-1] == 276 and cfg.DATASET.trainName[70] == 31
    assert not set(cfg.DATASET.trainName) & (set(cfg.DATASET.valName) | set(cfg.DATASET.testName))
    path = str(tmp_path / "cfg.yaml")
    config.dump_default_yaml(path)
    again = config.load_config(path)
    assert again.DATASET.trainName == cfg.DATASET.trainName and again.TRAINING.lr == 1e-4 and again.DATASET.idxToJoints[6] == "Neck"
    shipped = config.load_config(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config", "mscsa_prgcn.yaml"))
    assert shipped.DATASET.testName == cfg.DATASET.testName and shipped.TEST.batchSize == 32


def test_cli_flags_match_reference_main():
    from hupr_b200.main import build_parser
    args = build_parser().parse_args(["--config", "mscsa_prgcn.yaml", "--dir", "run1", "--eval", "-sr", "2", "--gpuIDs", "[0,1]", "--keypoints"])
    assert (args.dir, args.eval, args.sampling_ratio, args.gpuIDs, args.keypoints, args.seed, args.visDir) == ("run1", True, 2, [0, 1], True, 0, "none")


def test_window_index_map_equals_reference_loop():
    from hupr_b200.datasets import window_frame_indices
    for duration in (600, 12):
        for index in list(range(0, 3 * duration)):
            assert window_frame_indices(index, duration) == loader.window_indices(index, duration)


def test_gt_annotation_and_dataset_bookkeeping(tmp_path):
    from hupr_b200 import config
    from hupr_b200.datasets import HuPR3D_horivert, generateGTAnnot
    root = str(tmp_path / "data")
    blocks = make_dataset_dir(root, [7, 9], 3)
    with open(os.path.join(root, "hrnet_annot_train.json"), "w") as fp:
        json.dump(blocks, fp)
    cfg = config.default_config(DATASET__dataDir=root, DATASET__trainName=[7, 9], DATASET__duration=3)
    gt = json.load(open(generateGTAnnot(cfg, "train")))
    assert [im["id"] for im in gt["images"]] == [700000, 700001, 700002, 900000, 900001, 900002]
    a0 = gt["annotations"][0]
    assert a0["bbox"] == [10.0, 20.0, 190.0, 220.0] and a0["area"] == 190.0 * 220.0 / 2 and a0["num_keypoints"] == 14
    assert a0["keypoints"][2::3] == [2.0] * 14 and a0["keypoints"][0:2] == [float(v) for v in blocks[0][0]["joints"][0]]
    ds = HuPR3D_horivert("train", cfg, types.SimpleNamespace(sampling_ratio=1), random=False, device="cpu")
    assert len(ds) == 6 and ds.VRDAEPaths_hori[4].endswith("single_9/hori/000000001.npy") and ds.VRDAEPaths_vert[0].endswith("single_7/vert/000000000.npy")
    assert ds.annots[5]["imageId"] == 900002 and ds.annots[5]["joints"].shape == (14, 2)
    with pytest.raises(ValueError):
        HuPR3D_horivert("training", cfg, types.SimpleNamespace(sampling_ratio=1))


def test_pack_table_layout_of_the_batched_pack_launch():
    """ops.PackTable (host logic of hupr_pack_conv_weights_multi / hupr_unpack_wgrad_multi): CTA ranges of the jobs tile the grid without
    gaps, block_job names the owning job of every CTA, and the serialised jobs are the C struct's bytes."""
    import ctypes

    import torch
    from hupr_b200 import _C, ops
    jobs = [(0x1000, 0x2000, 0x3000, 0x4000, 0x5000, 64, 32, 27, 64, 64, 0),
            (0x6000, 0x7000, 0x8000, 0, 0, 14, 32, 1, 64, 64, 0),
            (0x9000, 0xa000, 0xb000, 0xc000, 0xd000, 128, 128, 9, 256, 128, 128)]
    t = ops.PackTable(jobs, torch.device("cpu"))
    blocks = [(-(-j[5] // 16)) * (-(-j[6] // 16)) for j in jobs]
    assert t.total_blocks == sum(blocks) and t.max_taps == 27
    assert t.block_job.tolist() == [0] * blocks[0] + [1] * blocks[1] + [2] * blocks[2]
    raw = bytes(t.jobs.numpy().tobytes())
    assert len(raw) == 3 * ctypes.sizeof(_C.PackJob)
    arr = (_C.PackJob * 3).from_buffer_copy(raw)
    assert [a.block_begin for a in arr] == [0, blocks[0], blocks[0] + blocks[1]]
    assert [(a.cout, a.cin, a.taps, a.cout_total, a.cin_pad, a.cout_off) for a in arr] == [j[5:] for j in jobs]
    assert arr[1].b_hi is None and arr[2].w == 0x9000
    import pytest
    with pytest.raises(ValueError):
        ops.PackTable([(1, 2, 3, 4, 5, 63, 32, 27, 64, 64, 0)], torch.device("cpu"))       # odd cout: bf16 pairs


def test_split_tensor_planes_are_halves_of_one_allocation():
    """ops.SplitTensor.empty / from_float: lo lies right above hi in ONE allocation (so that a tensor map can address both weight planes
    with a plane dimension: csrc/conv_halo.cu fetches hi and lo weight tiles with a single TMA box); odd sizes fall back to two
    allocations that keep both planes 128-byte aligned; the operand planes of the two-unit convolution have the documented shapes."""
    import torch
    from hupr_b200.ops import SplitTensor
    t = SplitTensor.empty((3, 4, 64), "cpu")
    assert t.lo.data_ptr() - t.hi.data_ptr() == t.hi.numel() * 2 and t.hi.is_contiguous() and t.lo.is_contiguous()
    x = torch.randn(2, 5, 64)
    f = SplitTensor.from_float(x)
    assert f.lo.data_ptr() - f.hi.data_ptr() == f.hi.numel() * 2
    assert torch.equal(f.hi, x.to(torch.bfloat16)) and torch.equal(f.lo, (x - x.to(torch.bfloat16).float()).to(torch.bfloat16))
    assert float((f.float() - x).abs().max()) <= float(x.abs().max()) * 2.0 ** -16
    odd = SplitTensor.empty((3, 5), "cpu")
    assert odd.hi.shape == odd.lo.shape == (3, 5)
    single = SplitTensor.from_float(x, split=False)
    assert single.lo is None and single.q is None and single.q_fresh is None
    q16, q8 = t.ensure_q()
    assert q16.shape == (3, 4, 64) and q16.dtype == torch.float16 and q8.shape == (3, 4, 128) and q8.dtype == torch.uint8
    import pytest
    with pytest.raises(ValueError):
        SplitTensor.empty((2, 48), "cpu").ensure_q()          # channel count must be a multiple of 32
