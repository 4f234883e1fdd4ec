"""Compact loader cache.

The reference stores every frame as a complex128 ``[16,64,64,8]`` cube (8 MiB; /root/reference/preprocessing/process_iwr1843.py:180-182)
although the loader reads Doppler rows 4..11 only and hands the network float32 (``/root/reference/datasets/dataset.py:139-150``,
``datasets/base.py:13-24``).  The plane cache holds exactly what ``HuPR3D_horivert.__getitem__`` builds per frame — the 8 kept rows,
(re, im), each (row, part, elevation) plane standardised over its 64x64 cells — as float32 ``[8,2,64,64,8]`` = 2 MiB, written by the
same GPU kernel the loader uses (``hupr_window_normalize``), so items built from either cache are bit-identical and the per-sample
CPU normalisation (1.25 s in the reference) disappears from the loader.  ``convert`` rewrites an existing cube cache in place of
re-running the FFT cascade.
"""
import os

import numpy as np
import torch

from .. import ops

PLANE_SHAPE = (8, 2, 64, 64, 8)


def planes_from_cubes(cubes, device="cuda"):
    """complex ``[n,16,64,64,8]`` cubes (numpy or tensor) -> float32 CUDA ``[n,8,2,64,64,8]`` standardised planes."""
    if isinstance(cubes, np.ndarray):
        cubes = torch.from_numpy(np.ascontiguousarray(cubes.astype(np.complex64)))
    cubes = cubes.to(device=device, dtype=torch.complex64).contiguous()
    n = cubes.shape[0]
    return ops.window_normalize(cubes, torch.arange(n, dtype=torch.int32, device=cubes.device))


def convert(src_dir, dst_dir, device="cuda", frames_per_launch=32):
    """Every ``*.npy`` cube under ``src_dir`` (any depth: ``single_<g>/{hori,vert}/%09d.npy``) -> same relative path under
    ``dst_dir`` in the plane format.  Returns the number of files written."""
    todo = []
    for root, _, files in os.walk(src_dir):
        for name in sorted(files):
            if name.endswith(".npy"):
                todo.append(os.path.relpath(os.path.join(root, name), src_dir))
    todo.sort()
    for i in range(0, len(todo), frames_per_launch):
        chunk = todo[i:i + frames_per_launch]
        cubes = []
        for rel in chunk:
            cube = np.load(os.path.join(src_dir, rel))
            if cube.shape != (16, 64, 64, 8):
                raise ValueError("%s: expected a [16,64,64,8] cube, got %s" % (rel, (cube.shape,)))
            cubes.append(cube.astype(np.complex64))
        planes = planes_from_cubes(np.stack(cubes), device).cpu().numpy()
        for rel, p in zip(chunk, planes):
            out = os.path.join(dst_dir, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            np.save(out, p)
    return len(todo)


if __name__ == "__main__":          # python -m hupr_b200.datasets.cubecache <cube cache dir> <plane cache dir>
    import sys
    if len(sys.argv) != 3:
        raise SystemExit("usage: python -m hupr_b200.datasets.cubecache <src dir of [16,64,64,8] .npy cubes> <dst dir>")
    print("%d files converted" % convert(sys.argv[1], sys.argv[2]))
