"""Python-side wrappers of the dense kernels (C ABI in include/hupr_b200.h).  PyTorch is used only for device
memory and streams; every arithmetic op here is a call into libhupr_b200.so."""
import os

import torch

import contextlib

from . import _C

_PROFILE = None     # list of [name, flops, start_event, stop_event] while a profile is being recorded


def profile_begin():
    """Start recording one (name, algorithmic FLOPs, CUDA-event duration) entry per library call (bench.py's breakdown)."""
    global _PROFILE
    _PROFILE = []


def profile_end():
    """Stop recording; returns ``[(name, flops, milliseconds), ...]`` in call order."""
    global _PROFILE
    rec, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    return [(name, flops, start.elapsed_time(stop)) for name, flops, start, stop in rec]


@contextlib.contextmanager
def _timed(name, flops=0):
    if _PROFILE is None:
        yield
        return
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    yield
    stop.record()
    _PROFILE.append([name, flops, start, stop])


class SplitTensor(object):
    """Channels-last activation ``[N, D, H, W, C]`` stored as two bf16 planes: value = hi + lo.

    ``lo is None`` means single-bf16 mode (fast, not fp32-equivalent).

    ``q`` (optional) holds the same values a second time as the operand planes of the two-unit 3-tap convolution
    (hupr_conv_desc.nprod == 2): ``(q16 float16 [..., C], q8 uint8 [..., 2 C])`` — q8 holds, per 32-channel block, 32 e4m3 values
    followed by the 32 e4m3 residuals.  ``q_fresh`` is the channel
    range ``(first, end)`` a producing convolution has just written there (one-shot: the consuming convolution clears it)."""

    __slots__ = ("hi", "lo", "q", "q_fresh")

    def __init__(self, hi, lo=None):
        self.hi = hi
        self.lo = lo
        self.q = None
        self.q_fresh = None

    def ensure_q(self):
        if self.q is None:
            if not self.hi.is_contiguous():
                raise ValueError("quantised operand planes need a dense tensor")
            dev, shape = self.hi.device, self.hi.shape
            if shape[-1] % 32:
                raise ValueError("quantised operand planes need a channel count that is a multiple of 32")
            self.q = (torch.empty(shape, dtype=torch.float16, device=dev),
                      torch.empty(tuple(shape[:-1]) + (2 * shape[-1],), dtype=torch.uint8, device=dev))
        return self.q

    @property
    def shape(self):
        return self.hi.shape

    @staticmethod
    def empty(shape, device, split=True, zero=False):
        """Both planes are halves of ONE allocation (lo right above hi): a tensor map can then address them as one tensor with a plane
        dimension, which lets the halo convolution fetch hi and lo weight tiles with a single TMA box (csrc/conv_halo.cu)."""
        mk = torch.zeros if zero else torch.empty
        if not split:
            return SplitTensor(mk(shape, dtype=torch.bfloat16, device=device), None)
        numel = 1
        for s in shape:
            numel *= int(s)
        if numel % 64:          # keep both planes 128-byte aligned
            return SplitTensor(mk(shape, dtype=torch.bfloat16, device=device), mk(shape, dtype=torch.bfloat16, device=device))
        both = mk((2,) + tuple(shape), dtype=torch.bfloat16, device=device)
        return SplitTensor(both[0], both[1])

    @staticmethod
    def from_float(x, split=True):
        """Host-side packing helper (weights, test inputs): fp32 -> hi/lo bf16."""
        x = x.float().contiguous()
        if not split:
            return SplitTensor(x.to(torch.bfloat16), None)
        out = SplitTensor.empty(x.shape, x.device, True)
        out.hi.copy_(x)
        out.lo.copy_(x - out.hi.float())
        return out

    def float(self):
        return self.hi.float() if self.lo is None else self.hi.float() + self.lo.float()


COOP_WS_BYTES = 40 << 20
_COOP_DEFAULT = os.environ.get("HUPR_NO_COOP", "") == ""      # A/B switch for measurements


class CoopWorkspaces(object):
    """Scratch buffers for the cooperative split-K of small-grid contractions (hupr_conv_desc.ws), owned by whoever owns the launch
    sequence (a HuPRNet plan, a TrainStep): one buffer per SLOT, where a slot is a set of launches that are ordered among themselves
    (slot 0 = the main stream, slot 1 = the side stream of the two-branch small-batch forward).  Buffers are allocated and zeroed
    EAGERLY; the kernels leave the arrival counters zero.  A slot that is first requested while a CUDA graph is being captured is an
    error: its zero-fill would only be recorded into that graph (and never run for other graphs sharing the buffer), so owners run an
    eager warm-up pass before capturing — which they need anyway for lazy weight packing."""

    def __init__(self, device):
        dev = torch.device(device)
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self._slots = {}

    def get(self, slot=0):
        ws = self._slots.get(slot)
        if ws is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("hupr_b200: split-K workspace slot %r first requested during CUDA-graph capture — run one eager "
                                   "warm-up pass of the same launch sequence before capturing" % (slot,))
            ws = torch.zeros(COOP_WS_BYTES, dtype=torch.uint8, device=self.device)
            self._slots[slot] = ws
        return ws


_COOP_GLOBAL = {}        # device index -> CoopWorkspaces used outside any owner scope, slots keyed by CUDA stream
_COOP_ACTIVE = None      # CoopWorkspaces of the owner whose launch sequence is running (coop_scope)
_COOP_SLOT = 0


@contextlib.contextmanager
def coop_scope(workspaces):
    """Route the split-K scratch of every conv_gemm call in the block to ``workspaces`` (a CoopWorkspaces owned by the caller)."""
    global _COOP_ACTIVE
    prev, _COOP_ACTIVE = _COOP_ACTIVE, workspaces
    try:
        yield workspaces
    finally:
        _COOP_ACTIVE = prev


@contextlib.contextmanager
def coop_slot(slot):
    """Launches in the block use scratch slot ``slot`` of the active scope (a branch that may run concurrently with slot 0)."""
    global _COOP_SLOT
    prev, _COOP_SLOT = _COOP_SLOT, slot
    try:
        yield
    finally:
        _COOP_SLOT = prev


def coop_workspace(device):
    """The scratch buffer for a conv_gemm launch on the current stream: the active owner scope's slot, else a process-wide buffer per
    (device, CUDA stream) — launches on one stream are ordered, launches on different streams must not share arrival counters."""
    if _COOP_ACTIVE is not None:
        return _COOP_ACTIVE.get(_COOP_SLOT)
    dev = torch.device(device)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    pool = _COOP_GLOBAL.get(index)
    if pool is None:
        pool = _COOP_GLOBAL[index] = CoopWorkspaces(torch.device("cuda", index))
    return pool.get(("stream", int(torch.cuda.current_stream(index).cuda_stream)))


_NPROD = 0      # tensor-core products per k-step requested by the running launch sequence (0 = by operands; see products())


@contextlib.contextmanager
def products(n):
    """Launches in the block use ``n`` tensor-core products per k-step in hupr_conv_gemm / hupr_conv_wgrad: 3 = fp32-equivalent hi/lo
    arithmetic (default when lo planes exist), 1 = plain bf16 products with fp32 accumulation (lo planes of the operands are not read)."""
    global _NPROD
    if n not in (0, 1, 3):
        raise ValueError("products must be 0, 1 or 3")
    prev, _NPROD = _NPROD, n
    try:
        yield
    finally:
        _NPROD = prev


_QUANT = False      # the running launch sequence asks for the two-unit arithmetic where a convolution's shape allows it (quant())
QUANT_FUSE = os.environ.get("HUPR_QUANT_FUSE", "1") != "0"      # A/B switch: producers write the planes vs a separate pass per consumer


@contextlib.contextmanager
def quant(on=True):
    """Convolutions in the block whose shape the two-unit kernel handles (3-tap, cout a multiple of 128, enough tiles) run with
    hupr_conv_desc.nprod == 2: fp16 main product + e4m3 cross terms (include/hupr_b200.h; 2/3 of the tensor-core work, error of the same
    order as the three bf16 products).  Inference only: the planes' fixed scales assume activation magnitudes (full accuracy for
    |x| < 224, hard limit 16 376; weights below 3.99 — checked when their planes are made), not gradients."""
    global _QUANT
    prev, _QUANT = _QUANT, bool(on)
    try:
        yield
    finally:
        _QUANT = prev


_QUANT_SAT = {}      # device index -> int32[1] counter of values that left the fp16 operand planes' range (quant_saturations)


def _quant_sat(device):
    dev = torch.device(device)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    t = _QUANT_SAT.get(index)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            return None                 # counters are created by an eager pass; a graph captured without one simply does not report
        t = _QUANT_SAT[index] = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", index))
    return t


def quant_saturations(device=None, reset=True):
    """How many values the two-unit convolutions' operand planes could not represent (|activation| >= 16 376) since the last reset, on
    ``device`` — 0 in normal operation; anything else means those forwards were wrong and ``HuPRNet.quant_cross_terms`` should be off
    for this model / input scaling.  Synchronises the device."""
    index = None if device is None else torch.device(device).index
    t = _QUANT_SAT.get(torch.cuda.current_device() if index is None else index)
    if t is None:
        return 0
    n = int(t.item())
    if reset and n:
        t.zero_()
    return n


def quantize_planes(t, ch_off=0, ch=None, is_weight=False):
    """hupr_quantize_planes: fill ``t.q`` (allocated on first use) for channels [ch_off, ch_off + ch) from the hi/lo planes."""
    q16, q8 = t.ensure_q()
    ld = t.hi.shape[-1]
    ch = ld - ch_off if ch is None else ch
    rows = t.hi.numel() // ld
    with torch.cuda.device(t.hi.device), _timed("quantize_planes"):
        _C.check(_C.lib().hupr_quantize_planes(t.hi.data_ptr(), _C.optr(t.lo), rows, ld, ch_off, ch, q16.data_ptr(), q8.data_ptr(),
                                               1 if is_weight else 0, None if is_weight else _C.optr(_quant_sat(t.hi.device)),
                                               _C.stream_ptr()), "hupr_quantize_planes")
    return t.q


def conv_gemm(a, cin, weight, cout, kernel=(1, 1, 1), pad=(0, 0, 0), a_ch_off=0, w_batched=False,
              scale=None, shift=None, slope=None, residual=None, r_ch_off=0,
              out=None, o_ch_off=0, out_f32=None, w_ld=0, w_ch_off=0, k_split=0, w_k_off=0, row_vec=None, row_mode=0, coop=True, tma_store=True,
              nprod=None, lcin=None, lcout=None, lrows=None, stats=None, out_q=False, probe=False):
    """out = act(scale * conv(a[..., a_ch_off:a_ch_off+cin], weight) + shift + residual)  — see hupr_conv_gemm.

    a        : SplitTensor [N, D, H, W, Ca]
    weight   : SplitTensor [taps | N, cout, cin]
    out      : SplitTensor [N, Dout, H, W, Co] (written at channel offset o_ch_off) and/or
    out_f32  : fp32 tensor [N, Dout, H, W, ld]
    lcin, lcout, lrows : LOGICAL contraction / output widths / GEMM rows of the reference layer when the call is zero-padded to the
                         kernel's 64-channel / 128-row granularity — only used for the algorithmic FLOP count of the profile.
    out_q    : also store ``out`` as quantised operand planes (``out.q``) for a following two-unit convolution (see quant())
    probe    : do not launch; return True when this call would run on the two-unit kernel (inside quant())
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
    desc.w_ld, desc.w_ch_off = w_ld, w_ch_off
    desc.k_split, desc.w_k_off = k_split, w_k_off
    if not tma_store or os.environ.get("HUPR_NO_TMA_STORE"):
        desc.no_tma_store = 1
    desc.nprod = _NPROD if nprod is None else nprod
    if stats is not None:            # float64 [2, >= cout]: sum v and sum v^2 per output channel are ADDED (fused BatchNorm batch statistics)
        if stats.dtype != torch.float64 or stats.dim() != 2 or stats.shape[0] != 2 or stats.stride(1) != 1:
            raise TypeError("conv_gemm: stats must be a float64 [2, C] tensor")
        desc.stats, desc.stats_ld = stats.data_ptr(), stats.stride(0)
    if k_split <= 1 and coop and _COOP_DEFAULT:     # small grids split their contraction cooperatively (deterministic ordered reduction)
        ws = coop_workspace(a.hi.device)
        desc.ws, desc.ws_bytes = ws.data_ptr(), ws.numel()
    if row_vec is not None:          # per-position vector: 1 = exp(acc - v), 2 = residual * (acc - v)  (attention backward)
        desc.row_vec, desc.row_mode = row_vec.data_ptr(), row_mode
    if a.hi.stride(0) != d * h * w * ca:        # overlapping sliding-window view over a frame stream
        if tuple(a.hi.stride()[1:]) != (h * w * ca, w * ca, ca, 1) or (a.lo is not None and a.lo.stride() != a.hi.stride()):
            raise ValueError("conv_gemm: only the sample stride of A may be non-dense")
        desc.a_n_stride = a.hi.stride(0)
    desc.scale, desc.shift, desc.slope = _C.optr(scale), _C.optr(shift), _C.optr(slope)
    if residual is not None:
        desc.r_hi, desc.r_lo = residual.hi.data_ptr(), _C.optr(residual.lo)
        desc.r_ld, desc.r_ch_off = residual.hi.shape[-1], r_ch_off
    if out is not None:
        desc.o_hi, desc.o_lo = out.hi.data_ptr(), _C.optr(out.lo)
        desc.o_ld, desc.o_ch_off = out.hi.shape[-1], o_ch_off
    if out_f32 is not None:
        desc.o_f32, desc.o_f32_ld = out_f32.data_ptr(), out_f32.shape[-1]
    use_q = False
    if _QUANT and kernel[1] == 3 and desc.nprod in (0, 3) and a.lo is not None and weight.lo is not None and stats is None:
        desc.nprod = 2
        use_q = _C.lib().hupr_conv_quant_eligible(desc) == 1
        if use_q and weight.q is None:
            if torch.cuda.is_current_stream_capturing():
                use_q = False           # weight planes are made by an eager warm-up pass, never inside a captured graph
            elif float(weight.hi.abs().max()) >= 3.99:
                weight.q = False        # outside the fp16 plane's range (w * 2^14): this filter keeps the three bf16 products
            else:
                quantize_planes(weight, is_weight=True)
        if weight.q is False:
            use_q = False
        if not use_q:
            desc.nprod = 3
    if probe:
        return use_q
    if use_q:
        span = (a_ch_off, a_ch_off + min(cin, ca - a_ch_off))      # multiples of 32 (eligibility)
        if a.q is None or a.q_fresh is None or a.q_fresh[0] > span[0] or a.q_fresh[1] < span[1]:
            quantize_planes(a, span[0], span[1] - span[0])
        a.q_fresh = None                # one-shot: whoever rewrites `a` must refresh the planes
        desc.a_q16, desc.a_q8 = (t.data_ptr() for t in a.q)
        desc.w_q16, desc.w_q8 = (t.data_ptr() for t in weight.q)
    if out_q and out is not None and QUANT_FUSE:
        desc.o_q16, desc.o_q8 = (t.data_ptr() for t in out.ensure_q())
        desc.q_sat = _C.optr(_quant_sat(out.hi.device))
        out.q_fresh = (o_ch_off, o_ch_off + cout)
    d_out = d + 2 * pad[0] - kernel[0] + 1
    rows = n * d_out * h * w if lrows is None else lrows
    flops = 2.0 * rows * (cout if lcout is None else lcout) * (min(cin, ca - a_ch_off) if lcin is None else lcin) * kernel[0] * kernel[1] * kernel[2]
    tag = "conv_gemm.%s%s" % ("attn" if w_batched else ("k%dx%dx%d" % tuple(kernel)), ".q" if use_q else "")
    with torch.cuda.device(a.hi.device), _timed(tag, flops):
        _C.check(_C.lib().hupr_conv_gemm(desc, _C.stream_ptr()), "hupr_conv_gemm")
    return out if out is not None else out_f32


def attention_fwd(q, q_off, k, k_off, vt, c, out, o_off, residual=None, r_off=0, lse=None):
    """Fused attention (hupr_attention_fwd): q, k SplitTensors ``[B, .., S, ld]``, vt ``[B, c, S]``, out ``[B, .., S, ld_o]``."""
    b = q.hi.shape[0]
    s = vt.hi.shape[-1]
    desc = _C.AttnDesc()
    desc.q_hi, desc.q_lo, desc.q_ld, desc.q_off = q.hi.data_ptr(), _C.optr(q.lo), q.hi.shape[-1], q_off
    desc.k_hi, desc.k_lo, desc.k_ld, desc.k_off = k.hi.data_ptr(), _C.optr(k.lo), k.hi.shape[-1], k_off
    desc.vt_hi, desc.vt_lo = vt.hi.data_ptr(), _C.optr(vt.lo)
    if residual is not None:
        desc.r_hi, desc.r_lo, desc.r_ld, desc.r_off = residual.hi.data_ptr(), _C.optr(residual.lo), residual.hi.shape[-1], r_off
    desc.o_hi, desc.o_lo, desc.o_ld, desc.o_off = out.hi.data_ptr(), _C.optr(out.lo), out.hi.shape[-1], o_off
    desc.batch, desc.s, desc.c = b, s, c
    desc.lse = _C.optr(lse)          # float32 [b, s]: log-sum-exp of every query row (training saves it for the backward)
    with torch.cuda.device(q.hi.device), _timed("attention_fwd", 4.0 * b * s * s * c):
        _C.check(_C.lib().hupr_attention_fwd(desc, _C.stream_ptr()), "hupr_attention_fwd")
    return out


def attention_bwd(q, q_off, k, k_off, v, v_off, do, do_off, lse, rowdot, dq, dq_off, dk, dk_off, dv, dv_off, c=64):
    """Fused attention backward (hupr_attention_bwd, c == 64): q, k, v, do SplitTensors ``[B, .., S, ld]``; lse, rowdot float32 ``[B, S]``;
    dq, dk, dv float32 ``[B, S, ld]`` — the three gradients are ADDED at their channel offsets."""
    b = q.hi.shape[0]
    s = lse.shape[-1]
    desc = _C.AttnBwdDesc()
    desc.q_hi, desc.q_lo, desc.q_ld, desc.q_off = q.hi.data_ptr(), _C.optr(q.lo), q.hi.shape[-1], q_off
    desc.k_hi, desc.k_lo, desc.k_ld, desc.k_off = k.hi.data_ptr(), _C.optr(k.lo), k.hi.shape[-1], k_off
    desc.v_hi, desc.v_lo, desc.v_ld, desc.v_off = v.hi.data_ptr(), _C.optr(v.lo), v.hi.shape[-1], v_off
    desc.do_hi, desc.do_lo, desc.do_ld, desc.do_off = do.hi.data_ptr(), _C.optr(do.lo), do.hi.shape[-1], do_off
    desc.lse, desc.rowdot = lse.data_ptr(), rowdot.data_ptr()
    desc.dq, desc.dq_ld, desc.dq_off = dq.data_ptr(), dq.shape[-1], dq_off
    desc.dk, desc.dk_ld, desc.dk_off = dk.data_ptr(), dk.shape[-1], dk_off
    desc.dv, desc.dv_ld, desc.dv_off = dv.data_ptr(), dv.shape[-1], dv_off
    desc.batch, desc.s, desc.c = b, s, c
    with torch.cuda.device(q.hi.device), _timed("attention_bwd", 10.0 * b * s * s * c):
        _C.check(_C.lib().hupr_attention_bwd(desc, _C.stream_ptr()), "hupr_attention_bwd")


@contextlib.contextmanager
def pdl(on=True):
    """Launches in the block use programmatic dependent launch (hupr_set_pdl): a kernel's launch + prologue overlap its predecessor's
    tail.  Worth it only in the latency regime (small-batch forwards); the setting is restored on exit."""
    prev = _C.lib().hupr_set_pdl(1 if on else 0)
    try:
        yield
    finally:
        _C.lib().hupr_set_pdl(prev)


def launch_count():
    """Kernels launched by libhupr_b200.so in this process so far (hupr_launch_count)."""
    return int(_C.lib().hupr_launch_count())


def _call(name, *args):
    with _timed(name[5:]):
        _C.check(getattr(_C.lib(), name)(*args), name)


def _p(t):
    return None if t is None else t.data_ptr()


def window_normalize(cube, slot_fs, out=None):
    """cube complex64 [n_fs,16,64,64,8], slot_fs int32 [n_slots] -> float32 [n_slots, 8, 2, 64, 64, 8] (hupr_window_normalize)."""
    if cube.dtype != torch.complex64 or not cube.is_contiguous() or slot_fs.dtype != torch.int32:
        raise TypeError("window_normalize expects a contiguous complex64 cube tensor and int32 slot indices")
    n_slots = slot_fs.numel()
    if out is None:
        out = torch.empty((n_slots, 8, 2, 64, 64, 8), dtype=torch.float32, device=cube.device)
    with torch.cuda.device(cube.device):
        _call("hupr_window_normalize", _p(cube), _p(slot_fs), n_slots, _p(out), _C.stream_ptr())
    return out


def plane_stats(cube, ws=None):
    """cube complex64 [n,16,64,64,8] -> per-plane mean / rstd workspace (hupr_plane_stats)."""
    n = cube.shape[0]
    if ws is None:
        ws = torch.empty(n * 256, dtype=torch.float32, device=cube.device)
    with torch.cuda.device(cube.device):
        _call("hupr_plane_stats", _p(cube), n, _p(ws), ws.numel() * 4, _C.stream_ptr())
    return ws


def frame_features(cube, stats, first, count, weight, bias, out):
    """Frame-sensors [first, first+count) of cube -> SplitTensor [count, 64, 64, 32] chirp features (hupr_frame_features)."""
    with torch.cuda.device(cube.device):
        _call("hupr_frame_features", _p(cube), _p(stats), first, count, _p(weight), _p(bias), _p(out.hi), _p(out.lo), _C.stream_ptr())
    return out


def mnet_fwd(vrdae, weight, bias, out):
    """vrdae float32 [..., 8, 2, 64, 64, 8] (n_slots leading elements) -> SplitTensor [n_slots, 64, 64, 32] (hupr_mnet_fwd)."""
    n_slots = vrdae.numel() // (8 * 2 * 64 * 64 * 8)
    with torch.cuda.device(vrdae.device):
        _call("hupr_mnet_fwd", _p(vrdae), _p(weight), _p(bias), _p(out.hi), _p(out.lo), n_slots, _C.stream_ptr())
    return out


def resample_linear(x, c, out, in_ch_off=0, out_ch_off=0):
    """align_corners=True bi/tri-linear resampling of channels [in_ch_off, +c) of x into channels [out_ch_off, +c) of out."""
    n, di, hi, wi, ild = x.hi.shape
    n2, do, ho, wo, old = out.hi.shape
    with torch.cuda.device(x.hi.device):
        _call("hupr_resample_linear", _p(x.hi), _p(x.lo), n, di, hi, wi, c, ild, in_ch_off,
              _p(out.hi), _p(out.lo), do, ho, wo, old, out_ch_off, _C.stream_ptr())
    return out


def softmax_rows(logits, out):
    """float32 [rows, cols] -> SplitTensor [rows, cols] (softmax along the last axis)."""
    cols = logits.shape[-1]
    rows = logits.numel() // cols
    with torch.cuda.device(logits.device):
        _call("hupr_softmax_rows", _p(logits), _p(out.hi), _p(out.lo), rows, cols, _C.stream_ptr())
    return out


def transpose_split(x, c, out, in_ch_off=0):
    """SplitTensor [n, .., s positions .., ld] channels [in_ch_off, +c) -> SplitTensor [n, c, s]."""
    n, ld = x.hi.shape[0], x.hi.shape[-1]
    s = x.hi.numel() // (n * ld)
    with torch.cuda.device(x.hi.device):
        _call("hupr_transpose_split", _p(x.hi), _p(x.lo), n, s, c, ld, in_ch_off, _p(out.hi), _p(out.lo), _C.stream_ptr())
    return out


def prgcn_workspace_bytes(batch):
    return int(_C.lib().hupr_prgcn_workspace_bytes(batch))


def prgcn_fwd(logits, weights, biases, adj, workspace, heatmap, gcn_heatmap):
    """logits float32 channels-last [B, 4096, ld] -> heatmap, gcn_heatmap float32 [B, 14, 64, 64] (hupr_prgcn_fwd)."""
    import ctypes
    batch, ld = logits.shape[0], logits.shape[-1]
    wp = (ctypes.c_void_p * 3)(*[w.data_ptr() for w in weights])
    bp = (ctypes.c_void_p * 3)(*[b.data_ptr() for b in biases])
    with torch.cuda.device(logits.device):
        _call("hupr_prgcn_fwd", _p(logits), ld, wp, bp, _p(adj), _p(workspace), workspace.numel() * workspace.element_size(),
              _p(heatmap), _p(gcn_heatmap), batch, _C.stream_ptr())
    return heatmap, gcn_heatmap


def gcn_nodes(logits, adj, heatmap, st):
    """logits float32 [B, 4096, ld] -> heatmap [B,14,64,64], st SplitTensor rows [(b,j)][1024] (hupr_gcn_nodes)."""
    batch, ld = logits.shape[0], logits.shape[-1]
    with torch.cuda.device(logits.device):
        _call("hupr_gcn_nodes", _p(logits), ld, _p(adj), _p(heatmap), _p(st.hi), _p(st.lo), batch, _C.stream_ptr())


def gcn_mix(yt, adj, st, batch):
    with torch.cuda.device(adj.device):
        _call("hupr_gcn_mix", _p(yt.hi), _p(yt.lo), _p(adj), _p(st.hi), _p(st.lo), batch, _C.stream_ptr())


def gcn_heads(y, gcn_heatmap, batch):
    with torch.cuda.device(y.device):
        _call("hupr_gcn_heads", _p(y), y.shape[-1], _p(gcn_heatmap), batch, _C.stream_ptr())


def keypoints_argmax(maps, preds=None, maxvals=None):
    """float32 [..., 64, 64] -> float32 [..., 2] (x, y) of the first maximum (hupr_keypoints_argmax)."""
    lead = maps.shape[:-2]
    n_maps = maps.numel() // 4096
    if preds is None:
        preds = torch.empty(tuple(lead) + (2,), dtype=torch.float32, device=maps.device)
    with torch.cuda.device(maps.device):
        _call("hupr_keypoints_argmax", _p(maps), n_maps, _p(preds), _p(maxvals), _C.stream_ptr())
    return preds


def heatmap_loss_fwd(heatmap, gcn_heatmap, joints, want_targets=False):
    """Returns (losses float32 [3] = {loss, loss2, loss1}, gt2d float32 [B,14,2], targets or None) — hupr_heatmap_loss_fwd."""
    batch = joints.shape[0]
    dev = heatmap.device
    joints = joints.to(device=dev, dtype=torch.int64).contiguous()
    ws = torch.empty(batch * 14 * 2, dtype=torch.float64, device=dev)
    losses = torch.empty(3, dtype=torch.float32, device=dev)
    gt2d = torch.empty((batch, 14, 2), dtype=torch.float32, device=dev)
    targets = torch.empty((batch, 14, 64, 64), dtype=torch.float32, device=dev) if want_targets else None
    with torch.cuda.device(dev):
        _call("hupr_heatmap_loss_fwd", _p(heatmap), _p(gcn_heatmap), _p(joints), batch, _p(ws), ws.numel() * 8,
              _p(losses), _p(targets), _p(gt2d), _C.stream_ptr())
    return losses, gt2d, targets


def heatmap_loss_bwd(heatmap, gcn_heatmap, joints, d_heat_logits, d_gcn_pre, weights=(1.0, 1.0)):
    """Gradient of weights[0]*loss1 + weights[1]*loss2 w.r.t. the pre-sigmoid logits (hupr_heatmap_loss_bwd).
    d_heat_logits float32 [B, 4096, ld] channels-last (14 channels written), d_gcn_pre float32 [B, 14, 64, 64]."""
    batch = joints.shape[0]
    joints = joints.to(device=heatmap.device, dtype=torch.int64).contiguous()
    with torch.cuda.device(heatmap.device):
        _call("hupr_heatmap_loss_bwd", _p(heatmap), _p(gcn_heatmap), _p(joints), batch, d_heat_logits.shape[-1],
              float(weights[0]), float(weights[1]), _p(d_heat_logits), _p(d_gcn_pre), _C.stream_ptr())
    return d_heat_logits, d_gcn_pre


def heatmap_bwd(heatmap, gcn_heatmap, g_heat, g_gcn, d_heat_logits, d_gcn_pre):
    """Sigmoid backward of both output maps for caller-supplied gradients (hupr_heatmap_bwd); g_heat / g_gcn float32 [B,14,64,64] or None."""
    batch = heatmap.shape[0]
    with torch.cuda.device(heatmap.device):
        _call("hupr_heatmap_bwd", _p(heatmap), _p(gcn_heatmap), _p(g_heat), _p(g_gcn), batch, d_heat_logits.shape[-1],
              _p(d_heat_logits), _p(d_gcn_pre), _C.stream_ptr())
    return d_heat_logits, d_gcn_pre


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, step_dev=None, lr_dev=None):
    """One Adam step with coupled L2 on flat float32 CUDA buffers (hupr_adam_step); defaults = reference tools/base.py:47.
    ``step_dev``: optional int32 device scalar holding the 1-based step count (used instead of ``step``; CUDA-graph friendly).
    ``lr_dev``: optional float32 device scalar holding the learning rate (used instead of ``lr``; lets a schedule act on a captured step)."""
    n = params.numel()
    if not (grads.numel() == exp_avg.numel() == exp_avg_sq.numel() == n):
        raise ValueError("adam_step: buffer sizes differ")
    with torch.cuda.device(params.device):
        _call("hupr_adam_step", _p(params), _p(grads), _p(exp_avg), _p(exp_avg_sq), n, lr, betas[0], betas[1], eps, weight_decay, step,
              _p(step_dev), _p(lr_dev), _C.stream_ptr())
    return params


def conv_wgrad_direct(x, x_ch_off, cin, dy, y_ch_off, cout, kernel, pad, out=None, batched=False, tag="conv_wgrad", nprod=None, lcin=None, lcout=None):
    """dW[tap, ci, co] of a stride-1 convolution straight from the channels-last tensors (hupr_conv_wgrad): x SplitTensor
    [n, d, h, w, cx], dy SplitTensor [n, d_out, h, w, cy] -> float32 [taps, cin, cout] (ACCUMULATED into ``out`` when given).
    ``batched``: no sum over the samples, out is [n, taps, cin, cout]."""
    n, d, h, w, cx = x.hi.shape
    taps = kernel[0] * kernel[1] * kernel[2]
    if out is None:
        out = torch.zeros(((n,) if batched else ()) + (taps, cin, cout), dtype=torch.float32, device=x.hi.device)
    desc = _C.WgradDesc()
    desc.x_hi, desc.x_lo = x.hi.data_ptr(), _C.optr(x.lo)
    desc.n, desc.d, desc.h, desc.w, desc.cx = n, d, h, w, cx
    desc.x_ch_off, desc.cin = x_ch_off, cin
    desc.dy_hi, desc.dy_lo, desc.cy = dy.hi.data_ptr(), _C.optr(dy.lo), dy.hi.shape[-1]
    desc.y_ch_off, desc.cout = y_ch_off, cout
    desc.kd, desc.kh, desc.kw = kernel
    desc.pd, desc.ph, desc.pw = pad
    desc.dw, desc.dw_ld = out.data_ptr(), out.shape[-1]
    if batched:
        desc.batched, desc.dw_batch_stride = 1, out.stride(0)
    desc.nprod = _NPROD if nprod is None else nprod
    d_out = d + 2 * pad[0] - kernel[0] + 1
    flops = 2.0 * n * d_out * h * w * (cout if lcout is None else lcout) * (cin if lcin is None else lcin) * taps
    with torch.cuda.device(x.hi.device), _timed(tag, flops):
        _C.check(_C.lib().hupr_conv_wgrad(desc, _C.stream_ptr()), "hupr_conv_wgrad")
    return out


def pack_conv_weights(w, fwd, cout_off=0, dgrad=None):
    """w float32 [cout, cin, *k] (a parameter in torch layout) -> rows cout_off.. of fwd SplitTensor [taps, cout_total, cin_pad] and, when
    given, columns cout_off.. of dgrad SplitTensor [taps, cin_pad, cout_total] (flipped taps) — hupr_pack_conv_weights."""
    cout, cin = w.shape[0], w.shape[1]
    taps = w.numel() // (cout * cin)
    t2, cout_total, cin_pad = fwd.hi.shape
    if t2 != taps or not w.is_contiguous() or w.dtype != torch.float32:
        raise ValueError("pack_conv_weights: weight %s does not match the packed layout %s" % (tuple(w.shape), tuple(fwd.hi.shape)))
    with torch.cuda.device(w.device):
        _call("hupr_pack_conv_weights", _p(w), cout, cin, taps, _p(fwd.hi), _p(fwd.lo), cout_total, cin_pad, cout_off,
              None if dgrad is None else _p(dgrad.hi), None if dgrad is None else _p(dgrad.lo), _C.stream_ptr())


def unpack_wgrad(acc, cout_off, dst):
    """acc float32 [taps, cin_pad, cout_total] (conv_wgrad_direct) -> dst float32 [cout, cin, *k] (torch layout) — hupr_unpack_wgrad."""
    taps, cin_pad, cout_total = acc.shape
    cout, cin = dst.shape[0], dst.shape[1]
    if dst.numel() != cout * cin * taps or not dst.is_contiguous():
        raise ValueError("unpack_wgrad: gradient %s does not match the accumulator %s" % (tuple(dst.shape), tuple(acc.shape)))
    with torch.cuda.device(acc.device):
        _call("hupr_unpack_wgrad", _p(acc), taps, cin_pad, cout_total, cout_off, _p(dst), cout, cin, _C.stream_ptr())


class PackTable(object):
    """Device-resident job table of the batched (un)pack launches (hupr_pack_conv_weights_multi / hupr_unpack_wgrad_multi): built once
    from ``(w, a_hi, a_lo, b_hi, b_lo, cout, cin, taps, cout_total, cin_pad, cout_off)`` tuples of raw pointers / sizes."""

    def __init__(self, jobs, device):
        import ctypes
        import numpy as np
        arr = (_C.PackJob * max(len(jobs), 1))()
        block_job, begin, self.max_taps = [], 0, 1
        for i, (w, a_hi, a_lo, b_hi, b_lo, cout, cin, taps, cout_total, cin_pad, cout_off) in enumerate(jobs):
            if cout % 2 or cin % 2 or cout_off % 2 or cin_pad % 2 or cout_total % 2 or not 0 < taps <= 32 or cout_off + cout > cout_total or cin > cin_pad:
                raise ValueError("PackTable: job %d violates the layout constraints of hupr_pack_conv_weights" % i)
            bx, by = -(-cout // 16), -(-cin // 16)
            j = arr[i]
            j.w, j.a_hi, j.a_lo, j.b_hi, j.b_lo = w, a_hi, a_lo, b_hi, b_lo
            j.cout, j.cin, j.taps, j.cout_total, j.cin_pad, j.cout_off = cout, cin, taps, cout_total, cin_pad, cout_off
            j.blocks_x, j.block_begin = bx, begin
            block_job.extend([i] * (bx * by))
            begin += bx * by
            self.max_taps = max(self.max_taps, taps)
        self.key = tuple(jobs)
        self.total_blocks = begin
        raw = np.frombuffer(ctypes.string_at(ctypes.addressof(arr), ctypes.sizeof(arr)), dtype=np.uint8).copy()
        self.jobs = torch.from_numpy(raw).to(device)
        self.block_job = torch.tensor(block_job if block_job else [0], dtype=torch.int32, device=device)

    def pack(self):
        with torch.cuda.device(self.jobs.device):
            _call("hupr_pack_conv_weights_multi", _p(self.jobs), _p(self.block_job), self.total_blocks, self.max_taps, _C.stream_ptr())

    def unpack(self):
        with torch.cuda.device(self.jobs.device):
            _call("hupr_unpack_wgrad_multi", _p(self.jobs), _p(self.block_job), self.total_blocks, self.max_taps, _C.stream_ptr())


def reduce_f64(src, groups, n, dst):
    """dst[g] = float(sum_i src[g*n + i]) — hupr_reduce_f64 (float64 partial sums -> float32 gradient storage)."""
    with torch.cuda.device(src.device):
        _call("hupr_reduce_f64", _p(src), groups, n, _p(dst), _C.stream_ptr())


def broadcast_f32(src, dst):
    with torch.cuda.device(dst.device):
        _call("hupr_broadcast_f32", _p(src), _p(dst), dst.numel(), _C.stream_ptr())


def gcn_bias_rows(bias, batch, rows):
    """bias float32 [1024, 14] -> rows SplitTensor [.., n_rows, 1024] (hupr_gcn_bias_rows)."""
    with torch.cuda.device(bias.device):
        _call("hupr_gcn_bias_rows", _p(bias), batch, rows.hi.shape[-2], _p(rows.hi), _p(rows.lo), _C.stream_ptr())


def bump_i32(counter):
    with torch.cuda.device(counter.device):
        _call("hupr_bump_i32", _p(counter), _C.stream_ptr())


def zero_(t):
    """Stream-ordered zero fill of a contiguous tensor through cudaMemsetAsync (hupr_memset_zero): a memset node under graph capture."""
    if not t.is_contiguous():
        raise ValueError("zero_: tensor must be contiguous")
    with torch.cuda.device(t.device):
        _call("hupr_memset_zero", _p(t), t.numel() * t.element_size(), _C.stream_ptr())
    return t


def zeros(shape, dtype, device):
    return zero_(torch.empty(shape, dtype=dtype, device=device))


def matmul_tn(a, b, b_ch_off, c, out):
    """out[s, m, :c] += sum_n a[s, n, m] * b[s, n, b_ch_off : b_ch_off + c] per sample s (hupr_conv_wgrad, batched): a SplitTensor
    [B, rows, cols] row-major, b SplitTensor [B, rows, ld]; out float32 [B, cols, c].  The attention backward's dK = dS^T Q and
    dV = P^T dO without transposed copies of the [S, S] matrices."""
    bsz, rows, cols = a.hi.shape
    av = SplitTensor(a.hi.view(bsz, 1, 1, rows, cols), a.lo.view(bsz, 1, 1, rows, cols))
    bv = SplitTensor(b.hi.view(bsz, 1, 1, rows, b.hi.shape[-1]), b.lo.view(bsz, 1, 1, rows, b.hi.shape[-1]))
    return conv_wgrad_direct(av, 0, cols, bv, b_ch_off, c, (1, 1, 1), (0, 0, 0), out=out.view(bsz, 1, cols, c), batched=True, tag="matmul_tn")


class KMajorGeometry(object):
    """Zero-padded linear position index shared by the two operands of a weight-gradient GEMM (see hupr_to_kmajor).  Wp is rounded
    up to a multiple of 8 so that the depth/height parts of a tap offset are 16-byte aligned (a TMA requirement); the +-1 shifts
    along W are realised as separately stored, pre-shifted copies of the X operand (one per kw tap)."""

    def __init__(self, n, d, h, w, pad):
        self.n, self.d, self.h, self.w = n, d, h, w
        self.pd, self.ph, self.pw = pad
        self.dp, self.hp = d + 2 * pad[0], h + 2 * pad[1]
        self.wp = -(-(w + 2 * pad[2]) // 8) * 8
        self.ppad = -(-(n * self.dp * self.hp * self.wp) // 64) * 64

    def tap_offset(self, kd, kh):
        return ((kd - self.pd) * self.hp + (kh - self.ph)) * self.wp


def to_kmajor(x, ch_off, c, geom, out, shift=0):
    """x SplitTensor [n, d, h, w, ld] channels [ch_off, +c) -> out SplitTensor [c, geom.ppad], written at P - shift (interior cells
    only; out is zero-filled once by its owner).  ``x`` may have depth 1 while geom.d > 1 (a depth-valid convolution's output
    placed at d = 0)."""
    n, d, h, w, ld = x.hi.shape
    with torch.cuda.device(x.hi.device):
        _call("hupr_to_kmajor", _p(x.hi), _p(x.lo), n, d, h, w, ld, ch_off, c, _p(out.hi), _p(out.lo),
              geom.dp, geom.hp, geom.wp, geom.pd, geom.ph, geom.pw, shift, geom.ppad, _C.stream_ptr())
    return out


def to_kmajor_multi(x, ch_off, c, geom, out, first_shift, n_copies, rows_per_copy):
    """One read of x, ``n_copies`` position-major copies: copy k goes to rows [k*rows_per_copy, +c) of ``out`` shifted by
    first_shift + k (hupr_to_kmajor_multi).  Needs c % 64 == 0 and an even W; callers fall back to to_kmajor otherwise."""
    n, d, h, w, ld = x.hi.shape
    with torch.cuda.device(x.hi.device):
        _call("hupr_to_kmajor_multi", _p(x.hi), _p(x.lo), n, d, h, w, ld, ch_off, c, _p(out.hi), _p(out.lo),
              geom.dp, geom.hp, geom.wp, geom.pd, geom.ph, geom.pw, first_shift, n_copies, rows_per_copy, geom.ppad, _C.stream_ptr())
    return out


def conv_wgrad(xt, cin, dyt, rows, geom, kernel, out):
    """Weight gradient of a stride-1 convolution from position-major operands (hupr_conv_gemm with k_split / w_k_off).

    xt  : SplitTensor [kernel[2] * cin, ppad]: input activations, row block kw pre-shifted by kw - pw (to_kmajor(shift=...)); the
          kw taps are one GEMM (N = kernel[2] * cin), so the output-gradient operand is read once per (kd, kh) instead of once per tap
    dyt : SplitTensor [rows, ppad]: output gradients in the same index space, rows = cout rounded up to a multiple of 128 (zero rows)
    out : float32 [kernel[0] * kernel[1], rows, kernel[2] * cin], zero-filled by the caller."""
    ppad = geom.ppad
    n = kernel[2] * cin
    a = SplitTensor(dyt.hi.view(1, 1, 1, rows, ppad), None if dyt.lo is None else dyt.lo.view(1, 1, 1, rows, ppad))
    w = SplitTensor(xt.hi.view(1, n, ppad), None if xt.lo is None else xt.lo.view(1, n, ppad))
    tiles = max(1, rows // 128) * max(1, n // (128 if n % 128 == 0 else 64))
    k_split = max(1, min(ppad // 64, (296 + tiles - 1) // tiles))
    g = 0
    for kd in range(kernel[0]):
        for kh in range(kernel[1]):
            conv_gemm(a, ppad, w, n, out_f32=out[g].view(1, 1, 1, rows, n), k_split=k_split, w_k_off=geom.tap_offset(kd, kh))
            g += 1
    return out
