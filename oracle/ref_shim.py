"""Loader for the UNMODIFIED reference (TEST INFRASTRUCTURE; build container only).

/root/reference is read-only and is NOT present on the GPU box, so nothing that runs under
``-m gpu``, ``smoke()`` or ``bench.py`` may import this module.  It exists to (a) pin the
oracle restatements in ``oracle/`` against the reference's own code and (b) generate the
golden fixtures committed under ``tests/golden/`` (``oracle/make_golden.py``).

No reference source is copied: the files are exec'd / imported from where they lie, with
the mechanical shims listed in SURVEY.md §8 c:
  1. ``'complex_'`` dtype strings (numpy 2 removed the alias)  -> ``'complex128'``
  2. ``plot_utils`` / ``matplotlib`` / ``pycocotools`` stub modules (not installed)
  3. ``torch.Tensor.cuda`` -> identity on a GPU-less host (layers.py:112 hard-codes .cuda())
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HUPR_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "preprocessing"))


def _stub(name, **attrs):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def load_radar_object():
    """Return the reference ``RadarObject`` class (process_iwr1843.py:8)."""
    path = os.path.join(REFERENCE_ROOT, "preprocessing", "process_iwr1843.py")
    with open(path, "r") as fp:
        text = fp.read()
    text = text.replace("'complex_'", "'complex128'")
    _stub("plot_utils", PlotMaps=lambda *a, **k: None, PlotHeatmaps=lambda *a, **k: None)
    namespace = {"__name__": "hupr_reference_preproc"}
    exec(compile(text, path, "exec"), namespace)
    return namespace["RadarObject"]


def _install_model_stubs():
    import numpy as np
    import torch
    plt = _stub("matplotlib.pyplot")
    _stub("matplotlib", pyplot=plt)
    coco = _stub("pycocotools.coco", COCO=object)
    cocoeval = _stub("pycocotools.cocoeval", COCOeval=object)
    _stub("pycocotools", coco=coco, cocoeval=cocoeval)
    if not hasattr(np, "float"):
        np.float = float
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


class Cfg(object):
    """YAML dict -> nested attribute object (what main.py:7-13 builds)."""

    def __init__(self, d):
        for k, v in d.items():
            setattr(self, k, Cfg(v) if isinstance(v, dict) else v)


def load_cfg():
    import yaml
    with open(os.path.join(REFERENCE_ROOT, "config", "mscsa_prgcn.yaml")) as fp:
        return Cfg(yaml.safe_load(fp))


def load_model_classes():
    """Return dict of reference model classes (models/*.py)."""
    _install_model_stubs()
    from models.networks import HuPRNet
    from models.chirp_networks import MNet
    from models.layers import Encoder3D, BasicBlock2D, BasicBlock3D, MultiScaleCrossSelfAttentionPRGCN
    from models.gcn_networks import PRGCN, GCN_layers
    return dict(HuPRNet=HuPRNet, MNet=MNet, Encoder3D=Encoder3D, BasicBlock2D=BasicBlock2D,
                BasicBlock3D=BasicBlock3D, Decoder=MultiScaleCrossSelfAttentionPRGCN,
                PRGCN=PRGCN, GCN_layers=GCN_layers)


def load_misc():
    """Return (LossComputer, generateTarget, get_max_preds) from misc/*.py."""
    _install_model_stubs()
    from misc.losses import LossComputer
    from misc.utils import generateTarget
    from misc.metrics import get_max_preds
    return LossComputer, generateTarget, get_max_preds


def load_normalize():
    """Return the reference ``Normalize`` transform class (datasets/base.py:13)."""
    _install_model_stubs()
    from datasets.base import Normalize
    return Normalize


def load_cocoeval():
    """Return the reference's vendored (COCO, COCOeval) classes (misc/coco.py, misc/cocoeval.py — the files the reference's README
    copies over pycocotools).  Shims: ``misc.mask`` (compiled RLE helpers, unused for keypoints), matplotlib submodules, ``np.float``."""
    import importlib
    import numpy as np
    _install_model_stubs()
    _stub("matplotlib.collections", PatchCollection=object)
    _stub("matplotlib.patches", Polygon=object)
    _stub("misc.mask")
    if not hasattr(np, "float"):
        np.float = float
    import misc  # noqa: F401  (package from /root/reference)
    coco = importlib.import_module("misc.coco")
    cocoeval = importlib.import_module("misc.cocoeval")
    return coco.COCO, cocoeval.COCOeval
