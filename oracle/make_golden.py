"""Generate tests/golden/* by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden [cascade|dca|loader|model|loss|all]

The fixtures pin the oracle restatements (oracle/*.py) to the reference's own code; they are small
(strided samples + float64 checksums) so they can live in git.  /root/reference is needed to run this
script, never to run the tests.
"""
import os
import sys
import tempfile

import numpy as np

from . import cascade, ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASCADE_CASES = [(0, 0), (3, 1)]           # (frame_idx, sensor) seeds of cascade.synth_frame
SAMPLE = (slice(None, None, 2), slice(None, None, 8), slice(None, None, 8), slice(None))


def cube_checksums(cube):
    cube = np.asarray(cube, dtype=np.complex128)
    return np.array([cube.real.sum(), cube.imag.sum(), (np.abs(cube) ** 2).sum(), np.abs(cube).max()],
                    dtype=np.float64)


def make_cascade():
    RadarObject = ref_shim.load_radar_object()
    ro = RadarObject()
    out = {}
    for frame_idx, sensor in CASCADE_CASES:
        frame = cascade.synth_frame(frame_idx, sensor)
        ref = ro.generateHeatmap(frame)
        key = "f%d_s%d" % (frame_idx, sensor)
        out[key + "_sample"] = ref[SAMPLE]
        out[key + "_checksums"] = cube_checksums(ref)
        out[key + "_dopplerplane_absmean"] = np.abs(ref).mean(axis=(1, 2, 3))
        print("cascade", key, ref.shape, out[key + "_checksums"])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "cascade_reference.npz"), **out)


def make_dca():
    RadarObject = ref_shim.load_radar_object()
    ro = RadarObject()
    rng = np.random.default_rng(7)
    n_frames = 2
    words = rng.integers(-32768, 32768, n_frames * cascade.FRAME_I16).astype(np.int16)
    with tempfile.TemporaryDirectory() as d:
        words.tofile(os.path.join(d, "adc_data.bin"))
        ref = ro.getadcDataFromDCA1000(d)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "dca1000_reference.npz"),
                        seed=np.array(7), n_frames=np.array(n_frames), shape=np.array(ref.shape),
                        sample=ref[:, ::37, ::5], real_sum=np.array(ref.real.sum()), imag_sum=np.array(ref.imag.sum()),
                        weighted=np.array((ref * np.arange(ref.size).reshape(ref.shape)).sum()))
    print("dca", ref.shape)


TARGETS = {"cascade": make_cascade, "dca": make_dca}


def main(argv):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    which = argv[1:] or ["all"]
    for name, fn in TARGETS.items():
        if "all" in which or name in which:
            fn()


if __name__ == "__main__":
    main(sys.argv)
