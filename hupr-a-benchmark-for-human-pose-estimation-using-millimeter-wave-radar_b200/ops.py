"""Python-side wrappers of the dense kernels (C ABI in include/hupr_b200.h).  PyTorch is used only for device
memory and streams; every arithmetic op here is a call into libhupr_b200.so."""
import torch

from . import _C


class SplitTensor(object):
    """Channels-last activation ``[N, D, H, W, C]`` stored as two bf16 planes: value = hi + lo.

    ``lo is None`` means single-bf16 mode (fast, not fp32-equivalent)."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo=None):
        self.hi = hi
        self.lo = lo

    @property
    def shape(self):
        return self.hi.shape

    @staticmethod
    def empty(shape, device, split=True, zero=False):
        mk = torch.zeros if zero else torch.empty
        hi = mk(shape, dtype=torch.bfloat16, device=device)
        lo = mk(shape, dtype=torch.bfloat16, device=device) if split else None
        return SplitTensor(hi, lo)

    @staticmethod
    def from_float(x, split=True):
        """Host-side packing helper (weights, test inputs): fp32 -> hi/lo bf16."""
        x = x.float().contiguous()
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16) if split else None
        return SplitTensor(hi, lo)

    def float(self):
        return self.hi.float() if self.lo is None else self.hi.float() + self.lo.float()


def conv_gemm(a, cin, weight, cout, kernel=(1, 1, 1), pad=(0, 0, 0), a_ch_off=0, w_batched=False,
              scale=None, shift=None, slope=None, residual=None, r_ch_off=0,
              out=None, o_ch_off=0, out_f32=None):
    """out = act(scale * conv(a[..., a_ch_off:a_ch_off+cin], weight) + shift + residual)  — see hupr_conv_gemm.

    a        : SplitTensor [N, D, H, W, Ca]
    weight   : SplitTensor [taps | N, cout, cin]
    out      : SplitTensor [N, Dout, H, W, Co] (written at channel offset o_ch_off) and/or
    out_f32  : fp32 tensor [N, Dout, H, W, ld]
    """
    n, d, h, w, ca = a.hi.shape
    desc = _C.ConvDesc()
    desc.a_hi, desc.a_lo = a.hi.data_ptr(), _C.optr(a.lo)
    desc.n, desc.d, desc.h, desc.w, desc.ca = n, d, h, w, ca
    desc.a_ch_off, desc.cin = a_ch_off, cin
    desc.w_hi, desc.w_lo = weight.hi.data_ptr(), _C.optr(weight.lo)
    desc.cout = cout
    desc.kd, desc.kh, desc.kw = kernel
    desc.pd, desc.ph, desc.pw = pad
    desc.w_batched = 1 if w_batched else 0
    desc.scale, desc.shift, desc.slope = _C.optr(scale), _C.optr(shift), _C.optr(slope)
    if residual is not None:
        desc.r_hi, desc.r_lo = residual.hi.data_ptr(), _C.optr(residual.lo)
        desc.r_ld, desc.r_ch_off = residual.hi.shape[-1], r_ch_off
    if out is not None:
        desc.o_hi, desc.o_lo = out.hi.data_ptr(), _C.optr(out.lo)
        desc.o_ld, desc.o_ch_off = out.hi.shape[-1], o_ch_off
    if out_f32 is not None:
        desc.o_f32, desc.o_f32_ld = out_f32.data_ptr(), out_f32.shape[-1]
    with torch.cuda.device(a.hi.device):
        _C.check(_C.lib().hupr_conv_gemm(desc, _C.stream_ptr()), "hupr_conv_gemm")
    return out if out is not None else out_f32
