"""CPU tests: the cascade oracle against the golden vectors minted from the reference itself."""
import os

import numpy as np

from oracle import cascade
from oracle.make_golden import CASCADE_CASES, SAMPLE, cube_checksums


def test_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "cascade_reference.npz"))
    for frame_idx, sensor in CASCADE_CASES:
        key = "f%d_s%d" % (frame_idx, sensor)
        cube = cascade.generate_heatmap(cascade.synth_frame(frame_idx, sensor))
        assert cube.shape == (16, 64, 64, 8) and cube.dtype == np.complex128
        # the closed form is bit-identical to the reference's loops in complex128
        assert np.array_equal(cube[SAMPLE], g[key + "_sample"])
        assert np.allclose(cube_checksums(cube), g[key + "_checksums"], rtol=1e-12, atol=0)


def test_looped_port_equals_vectorised():
    frame = cascade.synth_frame(11, 0)
    assert np.array_equal(cascade.generate_heatmap_looped(frame), cascade.generate_heatmap(frame))


def test_dca1000_closed_form_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "dca1000_reference.npz"))
    rng = np.random.default_rng(int(g["seed"]))
    words = rng.integers(-32768, 32768, int(g["n_frames"]) * cascade.FRAME_I16).astype(np.int16)
    adc = cascade.dca1000_to_complex(words)
    assert tuple(g["shape"]) == adc.shape
    assert np.array_equal(adc[:, ::37, ::5], g["sample"])
    assert adc.real.sum() == float(g["real_sum"]) and adc.imag.sum() == float(g["imag_sum"])
    assert (adc * np.arange(adc.size).reshape(adc.shape)).sum() == complex(g["weighted"])
    # round trip through the synthesiser used by the GPU tests
    assert np.array_equal(cascade.complex_to_dca1000(adc), words)


def test_index_maps_are_permutations():
    ele, az, dop, rng_ = cascade.output_index_maps()
    assert sorted(ele) == list(range(8)) and sorted(az) == list(range(64))
    assert list(dop) == list(range(56, 64)) + list(range(0, 8))
    assert list(rng_) == list(range(94, 30, -1))


def test_doppler_zero_plane_is_roundoff_only(golden_dir):
    # SURVEY.md §7 trap 1: out[8] is the clutter-removed DC Doppler bin -> pure round-off in the reference
    g = np.load(os.path.join(golden_dir, "cascade_reference.npz"))
    plane = g["f0_s0_dopplerplane_absmean"]
    assert plane[8] < 1e-6 and plane[7] > 1e3 and plane[9] > 1e3
