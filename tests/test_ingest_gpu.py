"""pipelined DCA1000 capture ingest (RadarObject.processRadarDataHoriVert / processCapture), the compact
plane cache and its converter.  Bars: files of the cube cache == the cascade output of the same words (bit-exact, complex64 widened to
complex128); plane cache == hupr_window_normalize of those cubes (bit-exact: same kernel); loader items from either cache bit-identical."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _capture_words(n_frames, seed):
    from oracle import cascade
    return np.concatenate([cascade.complex_to_dca1000(cascade.synth_frame(seed + f, seed & 1)) for f in range(n_frames)])


def test_pipelined_capture_ingest_writes_reference_layout(tmp_path, monkeypatch):
    from hupr_b200.preprocessing.process_iwr1843 import RadarObject, cascade_i16, FRAME_WORDS
    monkeypatch.chdir(tmp_path)
    n_frames = 7                                                  # 3 chunks of 3, 3, 1 frames: both stages re-used, ragged tail
    words = {}
    for sub, seed in (("hori", 10), ("vert", 21)):
        d = "raw_data/iwr1843/HuPR/single_1/%s" % sub
        os.makedirs(d)
        words[sub] = _capture_words(n_frames, seed)
        words[sub].tofile(os.path.join(d, "adc_data.bin"))
    ro = RadarObject(numGroup=1, framesPerLaunch=3)
    ro.numFrame = n_frames
    ro.saveDirNameGroup = [str(tmp_path / "data" / "single_1")]
    ro.processRadarDataHoriVert()
    for sub in ("hori", "vert"):
        ref = cascade_i16(torch.from_numpy(words[sub]).cuda().view(n_frames, FRAME_WORDS)).cpu().numpy()
        for f in range(n_frames):
            got = np.load(str(tmp_path / "data" / "single_1" / sub / ("%09d.npy" % f)))
            assert got.dtype == np.complex128 and got.shape == (16, 64, 64, 8)
            assert np.array_equal(got, ref[f].astype(np.complex128)), (sub, f)
    # the capture file holds the reference's de-interleave convention (closed form of getadcDataFromDCA1000)
    frame0 = ro.getadcDataFromDCA1000("raw_data/iwr1843/HuPR/single_1/hori")[:, :192, :]
    from oracle import cascade
    assert np.array_equal(frame0, cascade.synth_frame(10, 0))


def test_plane_cache_equals_loader_normalisation_and_converter(tmp_path):
    from hupr_b200 import ops
    from hupr_b200.datasets import cubecache
    from hupr_b200.preprocessing.process_iwr1843 import RadarObject, cascade_i16, FRAME_WORDS
    n_frames = 5
    words = _capture_words(n_frames, 33)
    cube_dir, plane_dir, conv_dir = str(tmp_path / "cube"), str(tmp_path / "planes"), str(tmp_path / "converted")
    for d in (cube_dir, plane_dir):
        os.makedirs(d)
    RadarObject(numGroup=0, framesPerLaunch=2).processCapture(words, cube_dir)
    RadarObject(numGroup=0, framesPerLaunch=2, cacheFormat="planes").processCapture(words, plane_dir)
    cubes = cascade_i16(torch.from_numpy(words).cuda().view(n_frames, FRAME_WORDS))
    ref = ops.window_normalize(cubes, torch.arange(n_frames, dtype=torch.int32, device="cuda")).cpu().numpy()
    for f in range(n_frames):
        p = np.load(os.path.join(plane_dir, "%09d.npy" % f))
        assert p.dtype == np.float32 and p.shape == (8, 2, 64, 64, 8)
        assert np.array_equal(p, ref[f])
        assert os.path.getsize(os.path.join(plane_dir, "%09d.npy" % f)) * 3.9 < os.path.getsize(os.path.join(cube_dir, "%09d.npy" % f))
    assert cubecache.convert(cube_dir, conv_dir, frames_per_launch=3) == n_frames
    for f in range(n_frames):
        assert np.array_equal(np.load(os.path.join(conv_dir, "%09d.npy" % f)), ref[f])
    with pytest.raises(ValueError):
        RadarObject(numGroup=0, cacheFormat="fp8")


def test_loader_items_identical_from_cube_and_plane_caches(tmp_path):
    import json
    import types
    from hupr_b200 import config
    from hupr_b200.datasets import cubecache, getDataset
    from tests.test_host_mirrors import make_dataset_dir
    root = str(tmp_path / "data")
    blocks = make_dataset_dir(root, [7], 4, seed=2)
    with open(os.path.join(root, "hrnet_annot_test.json"), "w") as fp:
        json.dump(blocks, fp)
    compact = str(tmp_path / "compact")
    cubecache.convert(root, compact)
    for name in ("hrnet_annot_test.json",):
        with open(os.path.join(root, name)) as src, open(os.path.join(compact, name), "w") as dst:
            dst.write(src.read())
    args = types.SimpleNamespace(sampling_ratio=1)
    items = []
    for d in (root, compact):
        cfg = config.default_config(DATASET__dataDir=d, DATASET__testName=[7], DATASET__duration=4)
        ds = getDataset("test", cfg, args, random=False)
        items.append([ds[i] for i in range(len(ds))])
    for a, b in zip(*items):
        assert torch.equal(a["VRDAEmap_hori"], b["VRDAEmap_hori"]) and torch.equal(a["VRDAEmap_vert"], b["VRDAEmap_vert"])
        assert a["imageId"] == b["imageId"]


def test_training_adc_ingest_equals_oracle_cascade_and_loader():
    """TrainStep.prefetch_adc / take_prefetched_adc: pinned int16 DCA1000 words of a sample's 8-frame window -> copy stream -> FFT cascade ->
    window standardisation -> the step's float32 VRDAE inputs.  Against the oracle chain (complex128 cascade restatement + fp64 Normalize,
    process_iwr1843.py:106-173 + datasets/base.py:13-24) on the signal chirp slots; slot 4 is the Doppler-0 round-off plane (unit variance
    by construction, checked as such)."""
    import numpy as np
    import torch
    from hupr_b200.models import HuPRNet
    from hupr_b200.training import TrainStep
    from oracle import cascade, loader
    from oracle import model as om
    from tests.test_model_gpu import make_cfg
    net = HuPRNet(make_cfg())
    net.load_state_dict(om.make_state_dict(0))
    step = TrainStep(net.cuda().train())
    frames = {s: [cascade.synth_frame(f, s) for f in range(8)] for s in (0, 1)}
    words = {s: torch.from_numpy(np.stack([cascade.complex_to_dca1000(f) for f in frames[s]])[None]).pin_memory() for s in (0, 1)}
    joints = torch.randint(0, 256, (1, 14, 2)).pin_memory()
    hori = torch.empty((1, 8, 8, 2, 64, 64, 8), device="cuda")
    vert = torch.empty_like(hori)
    jd = torch.empty((1, 14, 2), dtype=torch.int64, device="cuda")
    step.prefetch_adc(words[0], words[1], joints)
    step.take_prefetched_adc(hori, vert, jd)
    torch.cuda.synchronize()
    assert torch.equal(jd.cpu(), joints)
    signal = [c for c in range(8) if c != 4]
    for s, got in ((0, hori), (1, vert)):
        ref = loader.vrdae_from_cubes([cascade.generate_heatmap(f) for f in frames[s]])
        g = got[0].cpu().numpy()
        assert float(np.abs(g[:, signal] - ref[:, signal]).max()) < 2e-3
        noise = g[:, 4]
        assert np.isfinite(noise).all() and abs(float(noise.std()) - 1.0) < 0.05
