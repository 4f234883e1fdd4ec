"""Python wrappers of the training-mode kernels (csrc/train.cu; C ABI in include/hupr_b200.h).  A "view" argument is either a
``SplitTensor`` (all channels) or a ``(SplitTensor, ch_off)`` pair selecting channels ``[ch_off, ch_off + c)``."""
import ctypes

import torch

from . import _C
from .ops import SplitTensor, _call, _p

SUMS_STATS, SUMS_BN_BWD, SUMS_PRELU = 0, 1, 2


def _view(v):
    if v is None:
        return None
    t, off = (v, 0) if isinstance(v, SplitTensor) else v
    tv = _C.TensorView()
    tv.hi, tv.lo, tv.ld, tv.ch_off = t.hi.data_ptr(), _C.optr(t.lo), t.hi.shape[-1], off
    return ctypes.byref(tv)


def _positions(v):
    t = v if isinstance(v, SplitTensor) else v[0]
    return t.hi.numel() // t.hi.shape[-1]


def _dev(v):
    t = v if isinstance(v, SplitTensor) else v[0]
    return t.hi.device


def channel_sums(mode, a, c, s1, s2, b=None, mask=None, mean=None, rstd=None):
    with torch.cuda.device(_dev(a)):
        _call("hupr_channel_sums", mode, _view(a), _view(b), _view(mask), _p(mean), _p(rstd), _positions(a), c, _p(s1), _p(s2), _C.stream_ptr())


def affine_act(z, c, out, scale1=None, shift1=None, r=None, scale2=None, shift2=None, slope=None):
    with torch.cuda.device(_dev(z)):
        _call("hupr_affine_act", _view(z), _p(scale1), _p(shift1), _view(r), _p(scale2), _p(shift2), _p(slope), _view(out), _positions(z), c,
              _C.stream_ptr())


def bn_bwd_apply(g, z, c, mean, rstd, k1, k2, k3, out, mask=None):
    with torch.cuda.device(_dev(g)):
        _call("hupr_bn_bwd_apply", _view(g), _view(z), _view(mask), _p(mean), _p(rstd), _p(k1), _p(k2), _p(k3), _view(out), _positions(g), c,
              _C.stream_ptr())


def act_bwd(g, s, c, slope, out):
    with torch.cuda.device(_dev(g)):
        _call("hupr_act_bwd", _view(g), _view(s), _p(slope), _view(out), _positions(g), c, _C.stream_ptr())


def accumulate(out, c, a=None, b=None, f=None, f_off=0):
    with torch.cuda.device(_dev(out)):
        _call("hupr_accumulate", _view(a), _view(b), _p(f), 0 if f is None else f.shape[-1], f_off, _view(out), _positions(out), c, _C.stream_ptr())


def bn_finalize(sums, count, gamma, beta, eps, momentum, running_mean=None, running_var=None, num_batches_tracked=None):
    """sums: float64 [2, c] (sum z, sum z^2) -> float32 [4, c] rows (mean, rstd, scale, shift); running statistics updated in place
    (hupr_bn_finalize: one launch for the whole per-channel bookkeeping of a train-mode BatchNorm)."""
    c = sums[0].shape[-1]
    out = torch.empty((4, c), dtype=torch.float32, device=sums[0].device)
    with torch.cuda.device(sums[0].device):
        _call("hupr_bn_finalize", _p(sums[0]), _p(sums[1]), int(count), _p(gamma), _p(beta), float(eps), float(momentum), _p(out[0]), _p(out[1]),
              _p(out[2]), _p(out[3]), _p(running_mean), _p(running_var), _p(num_batches_tracked), c, _C.stream_ptr())
    return out


def bn_bwd_finalize(sums, count, dgamma=None, dbeta=None):
    """sums: float64 [2, c] (sum g', sum g' zhat) -> float32 [4, c] rows (k2, k3, dgamma, dbeta) (hupr_bn_bwd_finalize).  When ``dgamma`` /
    ``dbeta`` (contiguous float32 [c], e.g. the parameters' .grad views) are given, the two gradients are written there directly."""
    c = sums.shape[1]
    out = torch.empty((4, c), dtype=torch.float32, device=sums.device)
    with torch.cuda.device(sums.device):
        _call("hupr_bn_bwd_finalize", _p(sums[0]), _p(sums[1]), int(count), _p(out[0]), _p(out[1]), _p(out[2] if dgamma is None else dgamma),
              _p(out[3] if dbeta is None else dbeta), c, _C.stream_ptr())
    return out


def rowdot(a, b, c_n, out, sub=None):
    """out[pos] = sum_ch a[pos, ch] * (b[pos, ch] - sub[pos, ch]) (hupr_rowdot); a, b, sub are views, out float32 [positions]."""
    with torch.cuda.device(_dev(a)):
        _call("hupr_rowdot", _view(a), _view(b), _view(sub), _p(out), _positions(a), c_n, _C.stream_ptr())
    return out


def resample_linear_bwd(g, c, din, in_ch_off=0):
    """g: view of the forward OUTPUT gradient [n, do, ho, wo, ld]; din: float32 [n, di, hi, wi, ld_in] (+=, zero-filled by the caller)."""
    t = g if isinstance(g, SplitTensor) else g[0]
    n, do, ho, wo, _ = t.hi.shape
    _, di, hi, wi, ld = din.shape
    with torch.cuda.device(din.device):
        _call("hupr_resample_linear_bwd", _view(g), n, do, ho, wo, c, _p(din), di, hi, wi, ld, in_ch_off, _C.stream_ptr())


def softmax_bwd_rows(p, dp, ds):
    cols = dp.shape[-1]
    rows = dp.numel() // cols
    with torch.cuda.device(dp.device):
        _call("hupr_softmax_bwd_rows", _p(p.hi), _p(p.lo), _p(dp), _p(ds.hi), _p(ds.lo), rows, cols, _C.stream_ptr())
