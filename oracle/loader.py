"""Oracle (TEST INFRASTRUCTURE): restatement of the loader bridge between the cascade cubes and the network input.

Follows /root/reference/datasets:
  Normalize.__call__               base.py:13-24      per elevation channel over the 64x64 plane: min-max then standardise
                                                      (unbiased std), fp64 arithmetic (ToTensor keeps the ndarray's float64)
  HuPR3D_horivert.__getitem__      dataset.py:120-150 8-frame window clamped to the 600-frame capture, Doppler rows 4..11,
                                                      real/imag planes normalised separately -> float32 [8, 8, 2, 64, 64, 8]

Pinned against the reference's Normalize / window loop by tests/golden/loader_reference.npz (oracle/make_golden.py loader).
"""
import numpy as np

GROUP_FRAMES = 8
NUM_CHIRPS_KEPT = 8
NUM_DOPPLER = 16
DURATION = 600


def normalize_plane(x):
    """x: float64 [64, 64, 8] (range, azimuth, elevation) -> same shape, statistics over (range, azimuth) per elevation."""
    x = np.asarray(x, dtype=np.float64)
    z = x - x.min(axis=(0, 1), keepdims=True)
    z = z / z.max(axis=(0, 1), keepdims=True)
    mean = z.mean(axis=(0, 1), keepdims=True)
    std = z.std(axis=(0, 1), ddof=1, keepdims=True)
    return (z - mean) / std


def window_indices(index, duration=DURATION, group=GROUP_FRAMES):
    """The loop of dataset.py:126-138, verbatim control flow: capture-global frame index per window slot."""
    pad = index % duration
    idx = index - group // 2 - 1
    out = []
    for j in range(group):
        if j + pad <= group // 2:
            idx = index - pad
        elif j > (duration - 1 - pad) + group // 2:
            idx = index + (duration - 1 - pad)
        else:
            idx += 1
        out.append(idx)
    return out


def vrdae_from_cubes(cubes):
    """cubes: list of 8 complex [16, 64, 64, 8] cubes (one window) -> float32 [8, 8, 2, 64, 64, 8]."""
    out = np.zeros((GROUP_FRAMES, NUM_CHIRPS_KEPT, 2, 64, 64, 8), dtype=np.float32)
    lo = NUM_DOPPLER // 2 - NUM_CHIRPS_KEPT // 2
    for j, cube in enumerate(cubes):
        for s in range(NUM_CHIRPS_KEPT):
            out[j, s, 0] = normalize_plane(cube[lo + s].real)
            out[j, s, 1] = normalize_plane(cube[lo + s].imag)
    return out
