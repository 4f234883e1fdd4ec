"""GPU parity tests of the fused FFT cascade (through the C ABI) against the oracle and the golden vectors.

Tolerance (BASELINE.json north_star): 1e-3 relative on FFT magnitudes.  The clutter-removed DC Doppler plane
(out[8]) is pure round-off in the reference (SURVEY.md §7 trap 1), so it is checked against an absolute bound
of 1e-3 * max|cube| instead, and must not be exactly constant (downstream Normalize would produce NaN).
The integer index maps are checked bit-exactly with single-tone inputs whose peak location is an integer.
"""
import os

import numpy as np
import pytest
import torch

from oracle import cascade
from oracle.make_golden import CASCADE_CASES, SAMPLE

pytestmark = pytest.mark.gpu
REL_TOL = 1e-3


def run_gpu(words):
    from hupr_b200.preprocessing.process_iwr1843 import cascade_i16
    t = torch.from_numpy(np.ascontiguousarray(words)).cuda()
    out = cascade_i16(t)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def assert_cube_close(got, ref):
    scale = np.abs(ref).max()
    sig = [d for d in range(16) if d != 8]
    err = np.abs(got[sig] - ref[sig])
    # element-wise relative error on magnitudes, with the same absolute floor the survey used (1e-3 * max)
    assert err.max() <= REL_TOL * scale * 1e-2, "max abs err %g vs scale %g" % (err.max(), scale)
    mag_ref = np.abs(ref[sig])
    rel = np.abs(np.abs(got[sig]) - mag_ref) / np.maximum(mag_ref, 1e-30)
    assert np.quantile(rel, 0.999) <= REL_TOL
    assert rel[mag_ref > 1e-3 * scale].max() <= REL_TOL
    # DC Doppler plane: round-off only, but never exactly constant
    assert np.abs(got[8]).max() <= REL_TOL * scale
    assert np.isfinite(got).all()


def test_matches_oracle_on_synthetic_frames():
    frames = [cascade.synth_frame(i, s) for i in range(3) for s in (0, 1)]
    words = np.stack([cascade.complex_to_dca1000(f) for f in frames])
    got = run_gpu(words)
    assert got.shape == (6, 16, 64, 64, 8) and got.dtype == np.complex64
    for k, f in enumerate(frames):
        assert_cube_close(got[k], cascade.generate_heatmap(f))


def test_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "cascade_reference.npz"))
    for frame_idx, sensor in CASCADE_CASES:
        key = "f%d_s%d" % (frame_idx, sensor)
        got = run_gpu(cascade.synth_raw_i16(frame_idx, sensor)[None])[0]
        ref = g[key + "_sample"]
        scale = float(g[key + "_checksums"][3])
        sub = got[SAMPLE]
        keep = [i for i, d in enumerate(range(0, 16, 2)) if d != 8]
        assert np.abs(sub[keep] - ref[keep]).max() <= REL_TOL * 1e-2 * scale
        energy = (np.abs(got.astype(np.complex128)) ** 2).sum()
        assert abs(energy - g[key + "_checksums"][2]) <= 1e-4 * g[key + "_checksums"][2]


def test_index_maps_bit_exact_with_single_tones():
    """A unit-amplitude tone at integer (range, doppler) must peak at exactly the mapped (d, r) cell."""
    ele_src, az_src, dop_src, rng_src = cascade.output_index_maps()
    cases = [(31, 5), (94, -3), (62, 7), (47, -8), (80, 1)]
    words, expect = [], []
    samp = np.arange(256)[None, None, :]
    loop = (np.arange(192) // 3)[None, :, None]
    for rbin, dop in cases:
        frame = 1000.0 * np.exp(2j * np.pi * (rbin * samp / 256.0 + dop * loop / 64.0)) * np.ones((4, 1, 1))
        frame = np.rint(frame.real) + 1j * np.rint(frame.imag)
        words.append(cascade.complex_to_dca1000(frame))
        expect.append((int(np.where(dop_src == dop % 64)[0][0]), int(np.where(rng_src == rbin)[0][0])))
    got = run_gpu(np.stack(words))
    for k, (d, r) in enumerate(expect):
        power = (np.abs(got[k]) ** 2).sum(axis=(2, 3))
        assert np.unravel_index(power.argmax(), power.shape) == (d, r)
        ref = cascade.generate_heatmap(cascade.dca1000_to_complex(words[k]))
        pa = np.abs(got[k][d, r]); pr = np.abs(ref[d, r])
        assert np.unravel_index(pa.argmax(), pa.shape) == np.unravel_index(pr.argmax(), pr.shape)


def test_ragged_counts_and_edges():
    from hupr_b200.preprocessing.process_iwr1843 import cascade_i16, FRAME_WORDS
    # empty input
    empty = cascade_i16(torch.empty((0, FRAME_WORDS), dtype=torch.int16, device="cuda"))
    assert empty.shape == (0, 16, 64, 64, 8)
    # more frame-sensors than clusters, not a multiple of the cluster count: every slot must be written
    n = 151
    base = np.stack([cascade.synth_raw_i16(i, 0) for i in range(4)])
    words = base[np.arange(n) % 4]
    got = run_gpu(words)
    for k in range(n):
        assert np.array_equal(got[k], got[k % 4]), k
    assert_cube_close(got[3], cascade.generate_heatmap(cascade.synth_frame(3, 0)))
    # all-zero and full-scale inputs
    zeros = run_gpu(np.zeros((1, FRAME_WORDS), dtype=np.int16))
    assert not zeros.any()
    rng = np.random.default_rng(5)
    extreme = rng.choice(np.array([-32768, 32767], dtype=np.int16), size=(1, FRAME_WORDS))
    assert_cube_close(run_gpu(extreme)[0], cascade.generate_heatmap(cascade.dca1000_to_complex(extreme[0])))


def test_linearity_property_at_scale():
    """Size-independent property: cascade(a) + cascade(b) == cascade(a + b) (the whole path is linear)."""
    rng = np.random.default_rng(9)
    a = rng.integers(-8000, 8000, (2, cascade.FRAME_I16)).astype(np.int16)
    b = rng.integers(-8000, 8000, (2, cascade.FRAME_I16)).astype(np.int16)
    ga, gb, gab = run_gpu(a), run_gpu(b), run_gpu((a + b).astype(np.int16))
    scale = np.abs(gab).max()
    assert np.abs(ga + gb - gab).max() <= 1e-5 * scale


def test_bad_arguments_raise():
    from hupr_b200.preprocessing.process_iwr1843 import cascade_i16
    with pytest.raises(TypeError):
        cascade_i16(torch.zeros(10, dtype=torch.int16))
    with pytest.raises(ValueError):
        cascade_i16(torch.zeros(10, dtype=torch.int16, device="cuda"))


def test_radar_object_generate_heatmap_mirror():
    from hupr_b200.preprocessing.process_iwr1843 import RadarObject
    ro = RadarObject(numGroup=1)
    frame = cascade.synth_frame(2, 1)
    cube = ro.generateHeatmap(frame)
    assert cube.shape == (16, 64, 64, 8) and cube.dtype == np.complex128
    assert_cube_close(cube, cascade.generate_heatmap(frame))


def test_linearity_at_bench_size():
    """Size-independent property at the bench's full launch size (2048 frame-sensors = 1.5 GiB of words): every stage of the cascade
    (de-interleave, clutter removal, the three FFTs, crops / shifts / flips) is linear, so cascade(a + b) == cascade(a) + cascade(b)
    up to fp32 round-off.  No oracle involved: the CPU restatement needs seconds per frame."""
    from hupr_b200.preprocessing.process_iwr1843 import cascade_i16, FRAME_WORDS
    n = 2048
    gen = torch.Generator(device="cuda").manual_seed(77)
    a = torch.randint(-1000, 1000, (n, FRAME_WORDS), generator=gen, dtype=torch.int16, device="cuda")
    b = torch.randint(-1000, 1000, (n, FRAME_WORDS), generator=gen, dtype=torch.int16, device="cuda")
    out = torch.empty((n, 16, 64, 64, 8), dtype=torch.complex64, device="cuda")
    ref = cascade_i16(a).clone()
    ref += cascade_i16(b, out)
    got = cascade_i16(a + b, out)
    torch.cuda.synchronize()
    scale = float(torch.view_as_real(ref).abs().max())
    worst = 0.0
    for i in range(0, n, 256):                                     # chunked comparison keeps the temporaries small
        worst = max(worst, float(torch.view_as_real(got[i:i + 256] - ref[i:i + 256]).abs().max()))
    assert worst < 2e-5 * scale, (worst, scale)
    # and the launch is deterministic
    again = cascade_i16(a + b)
    assert torch.equal(again, got)
