"""Training step of MSCSA-PRGCN on the B200: train-mode forward (BatchNorm batch statistics), hand-assembled
backward pass and Adam, every arithmetic step a launch into libhupr_b200.so.

Replaces ``loss.backward(); optimizer.step()`` of /root/reference/tools/run.py:76-79 with ``optim.Adam(lr 1e-4, betas (.9,.999),
weight_decay 1e-4)`` of tools/base.py:47.  There is no autograd graph: the backward of each stage is written out against the
tensors the forward saved —
  convolutions : dgrad = the implicit-GEMM kernel with flipped filters; wgrad = hupr_conv_wgrad: positions contracted on the tensor
                 cores with MN-major operands read straight from the channels-last tensors — ops.conv_wgrad_direct
  BatchNorm    : batch statistics (double sums), affine + ReLU, standard three-term backward — train_ops
  attention    : head dim 64 (level 1, 94 % of the work): hupr_attention_bwd, the fused flash-style backward (no [S, S] matrix in memory);
                 other levels: P = exp(QK^T - lse) and dS = P o (dP - rowdot) rebuilt in GEMM epilogues, dQ = dS K, dK = dS^T Q, dV = P^T dO
                 (the two A^T B products contract over queries with MN-major operands — ops.matmul_tn, no transposed [S,S] copies)
  PRGCN        : the transposed-layout GEMMs of the forward with W^T, adjacency mixes with A^T, adjoint resampling
Inside a step NO framework kernel runs: the operand layouts of the tensor-core kernels are rebuilt from the flat fp32 parameter buffer by
hupr_pack_conv_weights, weight gradients land in the flat fp32 gradient buffer through hupr_unpack_wgrad / direct stores, and every
zero-initialised scratch tensor of the step comes out of ONE arena cleared by a single memset (``ZeroArena``).

Precision: ``products=3`` (default) keeps the fp32-equivalent hi/lo arithmetic everywhere; ``products=1`` runs every convolution, data
gradient and weight gradient as plain bf16 tensor-core products with fp32 accumulation (the "training bf16" of BASELINE.json configs[3]):
master weights, gradients and Adam moments stay fp32, BatchNorm / element-wise kernels and the attention kernels keep their hi/lo inputs.
"""
import contextlib
import os

import torch

from . import ops
from . import train_ops as T
from .models import layers as L
from .models.networks import ADJACENCY
from .ops import SplitTensor

EPS = 1e-5
MOMENTUM = 0.1
FUSED_BN_STATS = os.environ.get("HUPR_FUSED_BN_STATS", "1") != "0"      # A/B switch: 0 = separate hupr_channel_sums passes


# ------------------------------------------------------------------------------------------------------------------ scratch memory
class ZeroArena(object):
    """Zero-initialised scratch of ONE step for ONE batch size.  The first pass through the step ("planning") hands out individually
    cleared tensors and records their sizes; from the second pass on the same sequence of requests is served from one contiguous
    buffer that a single hupr_memset_zero clears at the start of the step."""

    def __init__(self, device):
        self.device = device
        self.sizes = []          # aligned byte sizes in request order
        self.buf = None
        self.cursor = 0
        self.offset = 0

    def begin(self):
        if self.buf is None and self.sizes and not torch.cuda.is_current_stream_capturing():
            self.buf = torch.empty(sum(self.sizes), dtype=torch.uint8, device=self.device)
        self.cursor, self.offset = 0, 0
        if self.buf is not None:
            ops.zero_(self.buf)
        else:
            self.sizes = []

    def take(self, shape, dtype):
        n = torch.empty((), dtype=dtype).element_size()
        for d in shape:
            n *= d
        aligned = -(-n // 256) * 256
        if self.buf is None:
            self.sizes.append(aligned)
            return ops.zeros(tuple(shape), dtype, self.device)
        if self.cursor >= len(self.sizes) or self.sizes[self.cursor] != aligned:
            cursor = self.cursor
            self.buf, self.sizes = None, []          # forget the plan: the next pass plans again instead of failing forever
            raise RuntimeError("hupr_b200.training: the step's scratch request sequence changed between passes (request %d: %d bytes)"
                               % (cursor, aligned))
        t = self.buf[self.offset:self.offset + n].view(dtype).view(tuple(shape))
        self.cursor += 1
        self.offset += aligned
        return t


_ARENA = None


@contextlib.contextmanager
def _arena_scope(arena):
    global _ARENA
    prev, _ARENA = _ARENA, arena
    try:
        if arena is not None:
            arena.begin()
        yield
    finally:
        _ARENA = prev


def _zeros(shape, dtype, dev):
    """Zero-initialised scratch for the running step: from the step's arena when one is active, else an individually cleared tensor."""
    if _ARENA is not None:
        return _ARENA.take(shape, dtype)
    return ops.zeros(tuple(shape), dtype, dev)


def _S(shape, dev):
    return SplitTensor.empty(tuple(shape), dev, True)


def _rows(t):
    """SplitTensor [n, d, h, w, c] -> view [n, 1, 1, d*h*w, c]."""
    n, c = t.hi.shape[0], t.hi.shape[-1]
    s = t.hi.numel() // (n * c)
    return SplitTensor(t.hi.view(n, 1, 1, s, c), t.lo.view(n, 1, 1, s, c))


_UNPACK_JOBS = "__unpack_jobs__"      # key under which TrainStep collects the (accumulator, gradient) pairs of its batched unpack launch
_KEEP_ALIVE = "__keep_alive__"


def _grad_out(grads, name, shape, dev):
    """Destination of a parameter gradient: the pre-installed tensor of ``grads`` (TrainStep: a view of the flat gradient buffer) or a
    fresh float32 tensor stored under ``name`` (block-level use with a plain dict)."""
    t = grads.get(name)
    if t is None:
        t = grads[name] = torch.empty(tuple(shape), dtype=torch.float32, device=dev)
    return t


class ConvOp(object):
    """One (possibly cout-concatenated) stride-1 convolution: forward / dgrad / wgrad operand packs and gradient bookkeeping."""

    def __init__(self, names, cin, couts, kernel, pad, bias_name=None):
        self.names, self.cin, self.couts = names, cin, couts
        self.kernel, self.pad, self.bias_name = tuple(kernel), tuple(pad), bias_name
        self.cin_pad = L.pad64(cin)
        self.cout_pads = [L.pad64(c) for c in couts]
        self.cout = sum(self.cout_pads)
        self.lcout = sum(couts)
        self.taps = kernel[0] * kernel[1] * kernel[2]
        self.w = self.wd = None
        self.shapes = {}

    def pack(self, sd):
        """sd: name -> float32 parameter (torch layout).  Rebuilds the forward [taps, cout, cin_pad] and data-gradient
        [flipped taps, cin_pad, cout] hi/lo operands in place (hupr_pack_conv_weights; the zero padding is written once, here)."""
        ws = self._bind(sd)
        o = 0
        for w, cp in zip(ws, self.cout_pads):
            ops.pack_conv_weights(w, self.w, o, self.wd)
            o += cp
        self._bind_bias(sd)

    def _bind(self, sd):
        """Allocate the (zero-padded) operand buffers on first use and record the parameter shapes; returns the fp32 sources."""
        ws = [sd[n].detach() for n in self.names]
        dev = ws[0].device
        if self.w is None or self.w.hi.device != dev:
            self.w = SplitTensor.empty((self.taps, self.cout, self.cin_pad), dev, True, zero=True)
            self.wd = SplitTensor.empty((self.taps, self.cin_pad, self.cout), dev, True, zero=True)
        out = []
        for n, w in zip(self.names, ws):
            if w.dtype != torch.float32 or not w.is_contiguous():
                w = w.float().contiguous()
            self.shapes[n] = tuple(w.shape)
            out.append(w)
        return out

    def pack_jobs(self, sd):
        """Job tuples of this convolution for the batched pack launch (ops.PackTable); the sources must be the LIVE fp32 parameters."""
        ws = self._bind(sd)
        self._bind_bias(sd)
        jobs, o = [], 0
        for w, cp in zip(ws, self.cout_pads):
            jobs.append((w.data_ptr(), self.w.hi.data_ptr(), self.w.lo.data_ptr(), self.wd.hi.data_ptr(), self.wd.lo.data_ptr(),
                         w.shape[0], w.shape[1], self.taps, self.cout, self.cin_pad, o))
            o += cp
        return jobs

    def _bind_bias(self, sd):
        b = sd[self.bias_name].detach() if self.bias_name else None
        self.bias = b if b is None or (b.dtype == torch.float32 and b.is_contiguous()) else b.float().contiguous()
        if self.bias_name:
            self.shapes[self.bias_name] = tuple(self.bias.shape)

    def forward(self, x, out, **kw):
        return ops.conv_gemm(x, self.cin_pad, self.w, self.cout, kernel=self.kernel, pad=self.pad, shift=self.bias, out=out,
                             lcin=self.cin, lcout=self.lcout, **kw)

    def dgrad(self, dy, out, **kw):
        """dy: SplitTensor [..., self.cout] (or a wider tensor with a_ch_off) -> out [..., cin_pad]."""
        dpad = (self.kernel[0] - 1 - self.pad[0], self.pad[1], self.pad[2])
        return ops.conv_gemm(dy, self.cout, self.wd, self.cin_pad, kernel=self.kernel, pad=dpad, out=out, lcin=self.lcout, lcout=self.cin, **kw)

    def wgrad(self, x, x_off, dy, dy_off, grads):
        """x: saved input [n, d, h, w, ld]; dy: output gradient [n, d_out, h, w, ld'] -> grads[name] (+ bias gradient).
        One hupr_conv_wgrad launch (positions contracted on the tensor cores straight from the channels-last tensors) into a zeroed
        [taps, cin_pad, cout] accumulator, then one hupr_unpack_wgrad per parameter into its torch-layout gradient."""
        dev = x.hi.device
        acc = _zeros((self.taps, self.cin_pad, self.cout), torch.float32, dev)
        ops.conv_wgrad_direct(x, x_off, self.cin_pad, dy, dy_off, self.cout, self.kernel, self.pad, out=acc, lcin=self.cin, lcout=self.lcout)
        o = 0
        deferred = grads.get(_UNPACK_JOBS)           # TrainStep: one batched unpack launch at the end of the backward pass
        for name, cp in zip(self.names, self.cout_pads):
            dst = _grad_out(grads, name, self.shapes[name], dev)
            if deferred is not None:
                deferred.append((acc.data_ptr(), dst.data_ptr(), 0, 0, 0, dst.shape[0], dst.shape[1], self.taps, self.cout, self.cin_pad, o))
                grads[_KEEP_ALIVE].append(acc)       # an individually allocated accumulator must outlive the deferred launch
            else:
                ops.unpack_wgrad(acc, o, dst)
            o += cp
        if self.bias_name:
            sums = _zeros((2, self.cout), torch.float64, dev)
            T.channel_sums(T.SUMS_PRELU, (dy, dy_off), self.cout, sums[0], sums[1])
            ops.reduce_f64(sums[1], self.couts[0], 1, _grad_out(grads, self.bias_name, self.shapes[self.bias_name], dev))


class BNOp(object):
    def __init__(self, prefix, c):
        self.prefix, self.c = prefix, c

    def forward_stats(self, z, z_off, params, buffers, count, sums=None):
        """Batch statistics of z[..., z_off:z_off+c]; updates running stats; returns (scale, shift) for the affine kernel.
        ``sums``: float64 [2, c] (sum z, sum z^2) already accumulated by the producing convolution's epilogue (hupr_conv_desc.stats);
        None = one hupr_channel_sums pass over z."""
        dev = z.hi.device
        if sums is None:
            sums = _zeros((2, self.c), torch.float64, dev)
            T.channel_sums(T.SUMS_STATS, (z, z_off), self.c, sums[0], sums[1])
        gamma, beta = params[self.prefix + ".weight"].detach(), params[self.prefix + ".bias"].detach()
        rm, rv = buffers[self.prefix + ".running_mean"], buffers[self.prefix + ".running_var"]
        nbt = buffers[self.prefix + ".num_batches_tracked"]
        ok = all(t.dtype == torch.float32 and t.is_contiguous() for t in (gamma, beta, rm, rv)) and nbt.dtype == torch.int64
        if ok:      # one launch: mean, rstd, fused affine, running statistics (hupr_bn_finalize)
            st = T.bn_finalize((sums[0], sums[1]), count, gamma, beta, EPS, MOMENTUM, rm, rv, nbt)
            self.mean, self.rstd, self.scale, self.shift = st[0], st[1], st[2], st[3]
            return self.scale, self.shift
        s1, s2 = sums[0], sums[1]          # parameters / buffers kept in another dtype (tests drive float64 state): same algebra in torch
        mean = s1 / count
        var = (s2 / count - mean * mean).clamp_min(0.0)
        self.mean, self.rstd = mean.float(), (1.0 / torch.sqrt(var + EPS)).float()
        gamma, beta = gamma.float(), beta.float()
        self.scale = (gamma * self.rstd).contiguous()
        self.shift = (beta - self.mean * self.scale).contiguous()
        rm.mul_(1 - MOMENTUM).add_(mean.float(), alpha=MOMENTUM)
        rv.mul_(1 - MOMENTUM).add_((var * (count / max(count - 1, 1))).float(), alpha=MOMENTUM)
        nbt.add_(1)
        return self.scale, self.shift

    def backward(self, g, z, z_off, mask, out, out_off, count, grads):
        """g: gradient w.r.t. the activation output (masked by mask > 0 when given) -> out[..., out_off:+c] = dz; fills dgamma, dbeta."""
        dev = z.hi.device
        sums = _zeros((2, self.c), torch.float64, dev)
        T.channel_sums(T.SUMS_BN_BWD, g, self.c, sums[0], sums[1], b=(z, z_off), mask=mask, mean=self.mean, rstd=self.rstd)
        st = T.bn_bwd_finalize(sums, count, dgamma=_grad_out(grads, self.prefix + ".weight", (self.c,), dev),
                               dbeta=_grad_out(grads, self.prefix + ".bias", (self.c,), dev))      # k2, k3, dgamma, dbeta in one launch
        T.bn_bwd_apply(g, (z, z_off), self.c, self.mean, self.rstd, self.scale, st[0], st[1], (out, out_off), mask=mask)


class Block3D(object):
    """BasicBlock3D (layers.py:40-70) in training mode."""

    def __init__(self, prefix, cin, cout):
        self.c = cout
        self.conv1 = ConvOp([prefix + ".main.0.weight", prefix + ".downsample.0.weight"], cin, [cout, cout], (3, 3, 3), (1, 1, 1))
        self.conv2 = ConvOp([prefix + ".main.3.weight"], cout, [cout], (3, 3, 3), (1, 1, 1))
        self.bn1, self.bn2, self.bnd = BNOp(prefix + ".main.1", cout), BNOp(prefix + ".main.4", cout), BNOp(prefix + ".downsample.1", cout)
        self.relu = None

    def pack(self, sd):
        self.conv1.pack(sd)
        self.conv2.pack(sd)

    def forward(self, x, params, buffers):
        dev, c = x.hi.device, self.c
        shape = x.hi.shape[:-1]
        count = x.hi.numel() // x.hi.shape[-1]
        if self.relu is None or self.relu.device != dev:
            self.relu = torch.zeros(c, dtype=torch.float32, device=dev)
        self.x = x
        # the batch statistics of the three BatchNorms ride in the epilogues of the convolutions that produce their inputs
        st1 = _zeros((2, 2 * c), torch.float64, dev) if FUSED_BN_STATS else None
        self.zc = self.conv1.forward(x, _S(shape + (2 * c,), dev), stats=st1)
        s1, h1 = self.bn1.forward_stats(self.zc, 0, params, buffers, count, sums=None if st1 is None else st1[:, :c])
        sd_, hd = self.bnd.forward_stats(self.zc, c, params, buffers, count, sums=None if st1 is None else st1[:, c:])
        self.t = _S(shape + (c,), dev)
        T.affine_act((self.zc, 0), c, self.t, scale1=s1, shift1=h1, slope=self.relu)
        st2 = _zeros((2, c), torch.float64, dev) if FUSED_BN_STATS else None
        self.z2 = self.conv2.forward(self.t, _S(shape + (c,), dev), stats=st2)
        s2, h2 = self.bn2.forward_stats(self.z2, 0, params, buffers, count, sums=st2)
        self.out = _S(shape + (c,), dev)
        T.affine_act(self.z2, c, self.out, scale1=s2, shift1=h2, r=(self.zc, c), scale2=sd_, shift2=hd, slope=self.relu)
        self.count = count
        return self.out

    def backward(self, dout, grads, need_dx=True):
        dev, c = dout.hi.device, self.c
        shape = dout.hi.shape[:-1]
        dzc = _S(shape + (2 * c,), dev)
        dz2 = _S(shape + (c,), dev)
        self.bn2.backward(dout, self.z2, 0, self.out, dz2, 0, self.count, grads)
        self.bnd.backward(dout, self.zc, c, self.out, dzc, c, self.count, grads)
        self.conv2.wgrad(self.t, 0, dz2, 0, grads)
        dt = self.conv2.dgrad(dz2, _S(shape + (c,), dev))
        self.bn1.backward(dt, self.zc, 0, self.t, dzc, 0, self.count, grads)
        self.conv1.wgrad(self.x, 0, dzc, 0, grads)
        if not need_dx:
            return None
        return self.conv1.dgrad(dzc, _S(shape + (self.conv1.cin_pad,), dev))


class Block2D(object):
    """BasicBlock2D without BatchNorm, PReLU activations (layers.py:24-38 as built by networks.py:21)."""

    def __init__(self, prefix, cin, cout):
        self.prefix, self.cp, self.cout = prefix, L.pad64(cout), cout
        self.conv1 = ConvOp([prefix + ".main.0.weight", prefix + ".downsample.0.weight"], cin, [cout, cout], (1, 3, 3), (0, 1, 1))
        self.conv2 = ConvOp([prefix + ".main.2.weight"], cout, [cout], (1, 3, 3), (0, 1, 1))
        self.a1 = self.a2 = None

    def pack(self, sd):
        self.conv1.pack(sd)
        self.conv2.pack(sd)
        self.pack_slopes(sd)

    def pack_slopes(self, sd):
        s1, s2 = sd[self.prefix + ".main.1.weight"].detach(), sd[self.prefix + ".relu.weight"].detach()
        dev = s1.device
        if self.a1 is None or self.a1.device != dev:
            self.a1 = torch.empty(self.cp, dtype=torch.float32, device=dev)
            self.a2 = torch.empty(self.cp, dtype=torch.float32, device=dev)
        # nn.PReLU's single slope -> the per-channel slope arrays of the kernels (hupr_broadcast_f32 reads the parameter in place)
        ops.broadcast_f32(s1 if s1.dtype == torch.float32 else s1.float(), self.a1)
        ops.broadcast_f32(s2 if s2.dtype == torch.float32 else s2.float(), self.a2)

    def forward(self, x, out=None, out_off=0):
        dev, cp = x.hi.device, self.cp
        shape = x.hi.shape[:-1]
        self.x = x
        self.zc = self.conv1.forward(x, _S(shape + (2 * cp,), dev))
        self.t = _S(shape + (cp,), dev)
        T.affine_act((self.zc, 0), cp, self.t, slope=self.a1)
        self.s = ops.conv_gemm(self.t, cp, self.conv2.w, cp, kernel=(1, 3, 3), pad=(0, 1, 1), residual=self.zc, r_ch_off=cp, out=_S(shape + (cp,), dev),
                               lcin=self.cout, lcout=self.cout)
        self.out = _S(shape + (cp,), dev) if out is None else out
        T.affine_act(self.s, cp, (self.out, out_off), slope=self.a2)
        return self.out

    def backward(self, dout, dout_off, grads):
        dev, cp = self.s.hi.device, self.cp
        shape = self.s.hi.shape[:-1]
        ds = _S(shape + (cp,), dev)
        T.act_bwd((dout, dout_off), self.s, cp, self.a2, ds)
        p = _zeros((2, cp), torch.float64, dev)
        T.channel_sums(T.SUMS_PRELU, (dout, dout_off), cp, p[0], p[1], b=self.s)
        ops.reduce_f64(p[0], 1, cp, _grad_out(grads, self.prefix + ".relu.weight", (1,), dev))
        self.conv2.wgrad(self.t, 0, ds, 0, grads)
        dt = self.conv2.dgrad(ds, _S(shape + (cp,), dev))
        dzc = _S(shape + (2 * cp,), dev)
        T.act_bwd(dt, (self.zc, 0), cp, self.a1, (dzc, 0))
        q = _zeros((2, cp), torch.float64, dev)
        T.channel_sums(T.SUMS_PRELU, dt, cp, q[0], q[1], b=(self.zc, 0))
        ops.reduce_f64(q[0], 1, cp, _grad_out(grads, self.prefix + ".main.1.weight", (1,), dev))
        T.accumulate((dzc, cp), cp, a=ds)
        self.conv1.wgrad(self.x, 0, dzc, 0, grads)
        return self.conv1.dgrad(dzc, _S(shape + (self.conv1.cin_pad,), dev))


class AttentionLevel(object):
    """One scale of the cross/self attention (layers.py:126-149) with its eight 1x1 projections."""

    def __init__(self, level, c, hw, prev):
        p = "radarDecoder."
        self.c, self.hw, self.s, self.prev = c, hw, hw * hw, prev
        self.fused_bwd = True        # hupr_attention_bwd where it applies (head dim 64); other levels keep the GEMM-epilogue formulation
        self.proj_h = ConvOp([p + "%s.%d.weight" % (n, level) for n in L.PROJ_HORI], c, [c] * 4, (1, 1, 1), (0, 0, 0))
        self.proj_v = ConvOp([p + "%s.%d.weight" % (n, level) for n in L.PROJ_VERT], c, [c] * 4, (1, 1, 1), (0, 0, 0))

    def pack(self, sd):
        self.proj_h.pack(sd)
        self.proj_v.pack(sd)

    def _plan(self):
        c, prev = self.c, self.prev
        # (q source, q offset, k source, k offset, v, output offset in cat, residual)
        return [("re", c, "ra", 0, "ra", prev, True), ("ra", 3 * c, "ra", 2 * c, "ra", prev + c, False),
                ("ra", c, "re", 0, "re", prev + 2 * c, True), ("re", 3 * c, "re", 2 * c, "re", prev + 3 * c, False)]

    def forward(self, ra, re, cat):
        dev, c, s = ra.hi.device, self.c, self.s
        b = ra.hi.shape[0]
        self.ra, self.re = _rows(ra), _rows(re)
        self.pr = {"ra": self.proj_h.forward(self.ra, _S((b, 1, 1, s, 4 * c), dev)), "re": self.proj_v.forward(self.re, _S((b, 1, 1, s, 4 * c), dev))}
        self.vt = {}
        for key, x in (("ra", self.ra), ("re", self.re)):
            self.vt[key] = ops.transpose_split(x, c, _S((b, c, s), dev))
        cat_s = _rows(cat)
        self.cat_s = cat_s
        self.lse = []                # per attention: float32 [b, s] log-sum-exp of the logit rows (fused forward only), else None
        v = {"ra": self.ra, "re": self.re}
        for qs, qo, ks, ko, vs, oo, res in self._plan():
            if c in (64, 128):
                lse = torch.empty((b, s), dtype=torch.float32, device=dev)
                ops.attention_fwd(self.pr[qs], qo, self.pr[ks], ko, self.vt[vs], c, cat_s, oo, residual=v[vs] if res else None, lse=lse)
                self.lse.append(lse)
            else:
                self.lse.append(None)
                logits = torch.empty((b, 1, 1, s, s), dtype=torch.float32, device=dev)
                kview = SplitTensor(self.pr[ks].hi.view(b, s, 4 * c), self.pr[ks].lo.view(b, s, 4 * c))
                ops.conv_gemm(self.pr[qs], c, kview, s, a_ch_off=qo, w_batched=True, w_ld=4 * c, w_ch_off=ko, out_f32=logits)
                probs = ops.softmax_rows(logits, _S((b, 1, 1, s, s), dev))
                ops.conv_gemm(probs, s, self.vt[vs], c, w_batched=True, residual=v[vs] if res else None, out=cat_s, o_ch_off=oo)
        return cat

    def backward(self, dcat, grads):
        """dcat: gradient of the concatenated maps [B, 1, hw, hw, prev + 4C] -> (d_ra, d_re) [B, 1, hw, hw, C]."""
        dev, c, s = dcat.hi.device, self.c, self.s
        b = dcat.hi.shape[0]
        dcat_s = _rows(dcat)
        v = {"ra": self.ra, "re": self.re}
        dpr = {"ra": _S((b, 1, 1, s, 4 * c), dev), "re": _S((b, 1, 1, s, 4 * c), dev)}
        # dV of both attentions that read a map as V accumulate in one fp32 buffer; the A^T B products (dK = dS^T Q, dV = P^T dO)
        # contract over the query axis with MN-major tensor-core operands (ops.matmul_tn): no transposed [S, S] copies
        dvf = {"ra": _zeros((b, s, c), torch.float32, dev), "re": _zeros((b, s, c), torch.float32, dev)}
        dv_res = {}
        if self.fused_bwd and c == 64 and all(l is not None for l in self.lse):
            # fused flash-style backward: no [S, S] matrix in memory; dQ / dK of the four attentions accumulate in fp32 projection-gradient
            # buffers (one conversion to hi/lo afterwards), dV in the per-source buffers
            dprf = {key: _zeros((b, s, 4 * c), torch.float32, dev) for key in ("ra", "re")}
            for idx, (qs, qo, ks, ko, vs, oo, res) in enumerate(self._plan()):
                rowdot = T.rowdot((dcat_s, oo), (self.cat_s, oo), c, torch.empty((b, s), dtype=torch.float32, device=dev),
                                  sub=v[vs] if res else None)
                ops.attention_bwd(self.pr[qs], qo, self.pr[ks], ko, v[vs], 0, dcat_s, oo, self.lse[idx], rowdot,
                                  dprf[qs], qo, dprf[ks], ko, dvf[vs], 0)
                if res:
                    dv_res[vs] = oo
            for key in ("ra", "re"):
                T.accumulate(dpr[key], 4 * c, f=dprf[key].view(b * s, 4 * c))
            plan = []
        else:
            plan = list(enumerate(self._plan()))
        for idx, (qs, qo, ks, ko, vs, oo, res) in plan:
            kview = SplitTensor(self.pr[ks].hi.view(b, s, 4 * c), self.pr[ks].lo.view(b, s, 4 * c))
            vview = SplitTensor(v[vs].hi.view(b, s, c), v[vs].lo.view(b, s, c))
            probs, dsm = _S((b, 1, 1, s, s), dev), _S((b, 1, 1, s, s), dev)
            if self.lse[idx] is not None:
                # P = exp(logits - lse) in the epilogue of the logits GEMM; dS = P * (dP - rowdot) in the epilogue of the dP GEMM, with
                # rowdot = sum_m P dP = <dO, P V> = <dO, O - residual> computed from the forward output: no fp32 [S, S] round trips
                ops.conv_gemm(self.pr[qs], c, kview, s, a_ch_off=qo, w_batched=True, w_ld=4 * c, w_ch_off=ko, out=probs,
                              row_vec=self.lse[idx], row_mode=1)
                rowdot = T.rowdot((dcat_s, oo), (self.cat_s, oo), c, torch.empty((b, s), dtype=torch.float32, device=dev),
                                  sub=v[vs] if res else None)
                ops.conv_gemm(dcat_s, c, vview, s, a_ch_off=oo, w_batched=True, out=dsm, residual=probs, row_vec=rowdot, row_mode=2)
            else:
                scratch = torch.empty((b, 1, 1, s, s), dtype=torch.float32, device=dev)
                ops.conv_gemm(self.pr[qs], c, kview, s, a_ch_off=qo, w_batched=True, w_ld=4 * c, w_ch_off=ko, out_f32=scratch)
                ops.softmax_rows(scratch, probs)
                ops.conv_gemm(dcat_s, c, vview, s, a_ch_off=oo, w_batched=True, out_f32=scratch)                   # dP = dO V^T
                T.softmax_bwd_rows(probs, scratch, dsm)
            kt = ops.transpose_split(self.pr[ks], c, _S((b, c, s), dev), in_ch_off=ko)
            ops.conv_gemm(dsm, s, kt, c, w_batched=True, out=dpr[qs], o_ch_off=qo)                                 # dQ = dS K
            dkf = _zeros((b, s, c), torch.float32, dev)
            ops.matmul_tn(SplitTensor(dsm.hi.view(b, s, s), dsm.lo.view(b, s, s)), SplitTensor(self.pr[qs].hi.view(b, s, 4 * c), self.pr[qs].lo.view(b, s, 4 * c)),
                          qo, c, dkf)                                                                              # dK = dS^T Q
            T.accumulate((dpr[ks], ko), c, f=dkf.view(b * s, c))
            ops.matmul_tn(SplitTensor(probs.hi.view(b, s, s), probs.lo.view(b, s, s)),
                          SplitTensor(dcat_s.hi.view(b, s, -1), dcat_s.lo.view(b, s, -1)), oo, c, dvf[vs])         # dV += P^T dO
            if res:
                dv_res[vs] = oo
        dv = {}
        for key in ("ra", "re"):
            dv[key] = _S((b, 1, 1, s, c), dev)
            T.accumulate(dv[key], c, b=(dcat_s, dv_res[key]) if key in dv_res else None, f=dvf[key].view(b * s, c))
        out = {}
        for key, proj in (("ra", self.proj_h), ("re", self.proj_v)):
            proj.wgrad(v[key], 0, dpr[key], 0, grads)
            dx = proj.dgrad(dpr[key], _S((b, 1, 1, s, c), dev), residual=dv[key])
            out[key] = SplitTensor(dx.hi.view(b, 1, self.hw, self.hw, c), dx.lo.view(b, 1, self.hw, self.hw, c))
        return out["ra"], out["re"]


class Encoder(object):
    def __init__(self, prefix, nf, g):
        self.prefix, self.nf, self.g = prefix, nf, g
        self.conv0 = ConvOp([prefix + ".layer1.0.weight"], nf, [2 * nf], (3, 3, 3), (1, 1, 1), bias_name=prefix + ".layer1.0.bias")
        self.blocks = [Block3D(prefix + ".layer1.1", 2 * nf, 2 * nf), Block3D(prefix + ".layer2.1", 2 * nf, 4 * nf),
                       Block3D(prefix + ".layer2.2", 4 * nf, 4 * nf), Block3D(prefix + ".layer3.1", 4 * nf, 8 * nf),
                       Block3D(prefix + ".layer3.2", 8 * nf, 8 * nf)]
        self.merges = [ConvOp([prefix + ".l1temporalMerge.weight"], 2 * nf, [2 * nf], (g, 1, 1), (0, 0, 0)),
                       ConvOp([prefix + ".l2temporalMerge.weight"], 4 * nf, [4 * nf], (g // 2, 1, 1), (0, 0, 0)),
                       ConvOp([prefix + ".temporalMerge.weight"], 8 * nf, [8 * nf], (g // 4, 1, 1), (0, 0, 0))]

    def pack(self, sd):
        self.conv0.pack(sd)
        for m in self.blocks + self.merges:
            m.pack(sd)

    def forward(self, x, params, buffers):
        dev, nf, g = x.hi.device, self.nf, self.g
        b = x.hi.shape[0]
        self.x0 = x
        self.l1a = self.conv0.forward(x, _S((b, g, 64, 64, 2 * nf), dev))
        self.l1 = self.blocks[0].forward(self.l1a, params, buffers)
        self.l2in = ops.resample_linear(self.l1, 2 * nf, _S((b, g // 2, 32, 32, 2 * nf), dev))
        self.l2 = self.blocks[2].forward(self.blocks[1].forward(self.l2in, params, buffers), params, buffers)
        self.l3in = ops.resample_linear(self.l2, 4 * nf, _S((b, g // 4, 16, 16, 4 * nf), dev))
        self.l3 = self.blocks[4].forward(self.blocks[3].forward(self.l3in, params, buffers), params, buffers)
        f1 = self.merges[0].forward(self.l1, _S((b, 1, 64, 64, 2 * nf), dev))
        f2 = self.merges[1].forward(self.l2, _S((b, 1, 32, 32, 4 * nf), dev))
        f3 = self.merges[2].forward(self.l3, _S((b, 1, 16, 16, 8 * nf), dev))
        return f1, f2, f3

    def backward(self, df1, df2, df3, grads):
        """Gradients of the three merged maps -> gradient of the chirp features [B, G, 64, 64, 64-padded]."""
        dev, nf, g = df1.hi.device, self.nf, self.g
        b = df1.hi.shape[0]
        self.merges[2].wgrad(self.l3, 0, df3, 0, grads)
        dl3 = self.merges[2].dgrad(df3, _S((b, g // 4, 16, 16, 8 * nf), dev))
        d = self.blocks[3].backward(self.blocks[4].backward(dl3, grads), grads)                 # -> d l3in [B, 2, 16, 16, 128]
        acc = _zeros((b, g // 2, 32, 32, 4 * nf), torch.float32, dev)
        T.resample_linear_bwd(d, 4 * nf, acc)
        self.merges[1].wgrad(self.l2, 0, df2, 0, grads)
        dl2 = self.merges[1].dgrad(df2, _S((b, g // 2, 32, 32, 4 * nf), dev))
        T.accumulate(dl2, 4 * nf, a=dl2, f=acc.view(-1, 4 * nf))
        d = self.blocks[1].backward(self.blocks[2].backward(dl2, grads), grads)                 # -> d l2in [B, 4, 32, 32, 64]
        acc = _zeros((b, g, 64, 64, 2 * nf), torch.float32, dev)
        T.resample_linear_bwd(d, 2 * nf, acc)
        self.merges[0].wgrad(self.l1, 0, df1, 0, grads)
        dl1 = self.merges[0].dgrad(df1, _S((b, g, 64, 64, 2 * nf), dev))
        T.accumulate(dl1, 2 * nf, a=dl1, f=acc.view(-1, 2 * nf))
        dl1a = self.blocks[0].backward(dl1, grads)
        self.conv0.wgrad(self.x0, 0, dl1a, 0, grads)
        return self.conv0.dgrad(dl1a, _S((b, g, 64, 64, L.pad64(nf)), dev))


class _Saved(object):
    """Tensors one forward pass leaves for its backward pass."""


class TrainStep(object):
    """``loss, loss2 = step.forward_backward(hori, vert, joints)`` fills ``param.grad`` of every model parameter;
    ``step.optimizer_step()`` applies Adam (coupled L2) in place.  ``model`` is a ``hupr_b200.models.HuPRNet`` on a CUDA device.

    ``products``: 3 = fp32-equivalent hi/lo tensor-core arithmetic (default), 1 = bf16 products with fp32 accumulation in every
    convolution / data-gradient / weight-gradient launch (fp32 master weights and Adam state either way).
    ``fused_attention_bwd``: hupr_attention_bwd for the head-dim-64 level (default on; ``HUPR_FUSED_ATTN_BWD=0`` or False selects the
    GEMM-epilogue formulation for A/B measurements)."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, fused_attention_bwd=None, products=3,
                 install_grads=True):
        if products not in (1, 3):
            raise ValueError("products must be 1 (bf16 products) or 3 (fp32-equivalent hi/lo products)")
        self.model = model
        self.products = products
        nf, g, kp = model.numFilters, model.numGroupFrames, model.numKeypoints
        self.nf, self.g, self.kp = nf, g, kp
        self.enc = {"ra": Encoder("RAradarEncoder", nf, g), "re": Encoder("REradarEncoder", nf, g)}
        self.levels = [AttentionLevel(0, 8 * nf, 16, 0), AttentionLevel(1, 4 * nf, 32, 4 * nf), AttentionLevel(2, 2 * nf, 64, 2 * nf)]
        if fused_attention_bwd is None:
            fused_attention_bwd = os.environ.get("HUPR_FUSED_ATTN_BWD", "1") != "0"
        for level in self.levels:
            level.fused_bwd = bool(fused_attention_bwd)
        p = "radarDecoder."
        self.dblocks = [Block2D(p + "decoderLayer3.0", 32 * nf, 8 * nf), Block2D(p + "decoderLayer3.1", 8 * nf, 4 * nf),
                        Block2D(p + "decoderLayer2.0", 20 * nf, 4 * nf), Block2D(p + "decoderLayer2.1", 4 * nf, 2 * nf),
                        Block2D(p + "decoderLayer1.0", 10 * nf, 2 * nf), Block2D(p + "decoderLayer1.1", 2 * nf, nf)]
        self.head = ConvOp([p + "decoderLayer1.2.weight"], nf, [kp], (1, 1, 1), (0, 0, 0))
        self.hyper = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.step_count = 0
        # flat parameter / gradient / moment buffers: parameters become views of one fp32 buffer so Adam is a single launch
        params = [q for q in model.parameters()]
        dev = params[0].device
        self.device = dev
        # every parameter starts on a 16-byte boundary of the flat buffers (the 1-element PReLU slopes would otherwise misalign their
        # successors; GEMM kernels store straight into gradient views); the few padding elements stay zero under Adam
        self.offsets, n = {}, 0
        for name, q in model.named_parameters():
            self.offsets[name] = n
            n += -(-q.numel() // 4) * 4
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.gview = {}
        for name, q in model.named_parameters():
            o, k = self.offsets[name], q.numel()
            self.flat_p[o:o + k].copy_(q.data.reshape(-1))
            q.data = self.flat_p[o:o + k].view(q.shape)
            self.gview[name] = self.flat_g[o:o + k].view(q.shape)
            if install_grads:
                q.grad = self.gview[name]
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.adj = torch.tensor(ADJACENCY, dtype=torch.float32, device=dev)
        self.adj_t = self.adj.t().contiguous()
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)      # read by hupr_adam_step: a schedule acts on captured steps
        self._lr_uploaded = float(lr)
        self.gcn_w = [SplitTensor.empty((1, 1024, 1024), dev, True, zero=True) for _ in range(3)]       # W  [q][p]: forward operand
        self.gcn_wt = [SplitTensor.empty((1, 1024, 1024), dev, True, zero=True) for _ in range(3)]      # W^T: operand of dSt = dYt W
        self.gcn_relu = torch.zeros(1024, dtype=torch.float32, device=dev)
        self.coop = ops.CoopWorkspaces(dev)       # split-K scratch of this step's launch sequence
        self._arenas = {}                         # batch size -> ZeroArena
        self._gcn_bufs = {}                       # batch size -> persistent zero-padded GCN row buffers
        self._dirty = True
        self._saved = None
        self._pack_table = None                   # batched pack launch (built on the first _pack)
        self._unpack_tables = {}                  # batch size -> (job key, ops.PackTable) of the batched gradient unpack
        self._gptr = {id(q): self.gview[name].data_ptr() for name, q in model.named_parameters()}
        model._train_step = self                  # one TrainStep owns a model's flat parameter storage (HuPRNet._forward_train reuses it)

    def gview_ptr(self, q):
        return self._gptr.get(id(q), 0)

    def _conv_ops(self):
        ops_ = []
        for e in self.enc.values():
            ops_.append(e.conv0)
            for blk in e.blocks:
                ops_ += [blk.conv1, blk.conv2]
            ops_ += e.merges
        for lvl in self.levels:
            ops_ += [lvl.proj_h, lvl.proj_v]
        for blk in self.dblocks:
            ops_ += [blk.conv1, blk.conv2]
        return ops_ + [self.head]

    # ---------------------------------------------------------------------------------------------------------------- packing
    def _pack(self):
        """Operand layouts of every contraction from the flat fp32 parameters — library launches only (hupr_pack_conv_weights,
        hupr_broadcast_f32), so a captured step re-packs after its own Adam launch without any framework kernel."""
        sd = {k: v.data for k, v in self.model.named_parameters()}
        for name in ("RAchirpNet.temporalConvWx1x1.weight", "radarDecoder.gcn.L3.bias"):      # first and a late parameter of the flat buffer
            if sd[name].data_ptr() != self.flat_p.data_ptr() + 4 * self.offsets[name]:
                raise RuntimeError("hupr_b200.training: the model's parameter storage was replaced after this TrainStep was built "
                                   "(model.to() / .float() / load into new tensors): build a new TrainStep")
        p = "radarDecoder."
        if self._pack_table is None:
            # ONE launch re-packs all 84 convolution filters and the three GCN matrices (hupr_pack_conv_weights_multi): the table holds raw
            # pointers into the flat parameter buffer and the persistent operand buffers, both fixed for the life of this object
            jobs = []
            for conv in self._conv_ops():
                jobs += conv.pack_jobs(sd)
            for i in range(3):
                w = sd[p + "gcn.L%d.weight" % (i + 1)]
                jobs.append((w.data_ptr(), self.gcn_w[i].hi.data_ptr(), self.gcn_w[i].lo.data_ptr(), self.gcn_wt[i].hi.data_ptr(),
                             self.gcn_wt[i].lo.data_ptr(), 1024, 1024, 1, 1024, 1024, 0))
            self._pack_table = ops.PackTable(jobs, self.device)
        self._pack_table.pack()
        for blk in self.dblocks:
            blk.pack_slopes(sd)
        self.gcn_b = [sd[p + "gcn.L%d.bias" % i] for i in (1, 2, 3)]
        self.mnet = {k: (sd[n + ".temporalConvWx1x1.weight"], sd[n + ".temporalConvWx1x1.bias"]) for k, n in (("ra", "RAchirpNet"), ("re", "REchirpNet"))}
        self._dirty = False

    def _gcn_buffers(self, b):
        bufs = self._gcn_bufs.get(b)
        if bufs is None:
            dev = self.device
            rows = -(-(b * self.kp) // 128) * 128
            mk = lambda: SplitTensor.empty((1, 1, 1, rows, 1024), dev, True, zero=True)     # padding rows stay zero: kernels write sample rows only
            bufs = dict(rows=rows, st=[mk() for _ in range(3)], dy=[mk() for _ in range(2)], bias=[mk() for _ in range(3)])
            self._gcn_bufs[b] = bufs
        return bufs

    def set_lr(self, lr):
        self.hyper["lr"] = float(lr)

    def _sync_lr(self):
        if self.hyper["lr"] != self._lr_uploaded:       # host-side schedule (Runner.adjustLR) -> device scalar, outside any captured graph
            self.lr_dev.fill_(float(self.hyper["lr"]))
            self._lr_uploaded = float(self.hyper["lr"])

    @contextlib.contextmanager
    def _scope(self, batch, arena=True):
        a = None
        if arena:
            a = self._arenas.get(batch)
            if a is None:
                a = self._arenas[batch] = ZeroArena(self.device)
        with torch.cuda.device(self.device), ops.coop_scope(self.coop), ops.products(self.products), _arena_scope(a):
            yield

    # ---------------------------------------------------------------------------------------------------------------- forward
    def _forward(self, hori, vert):
        """Train-mode forward; returns (heatmap, gcn_heatmap) float32 [B,14,64,64] and keeps what the backward needs in self._saved."""
        if self._dirty:
            self._pack()
        model, nf, g, kp = self.model, self.nf, self.g, self.kp
        dev = hori.device
        b = hori.shape[0]
        params = dict(model.named_parameters())
        buffers = dict(model.named_buffers())
        sv = self._saved = _Saved()
        sv.hori, sv.vert, sv.b = hori, vert, b
        chirp = {}
        for key, x in (("ra", hori), ("re", vert)):
            chirp[key] = ops.mnet_fwd(x.contiguous(), self.mnet[key][0], self.mnet[key][1], _S((b, g, 64, 64, nf), dev))
        feats = {key: self.enc[key].forward(chirp[key], params, buffers) for key in ("ra", "re")}
        l3, l2, l1 = self.levels
        cat3 = l3.forward(feats["ra"][2], feats["re"][2], _S((b, 1, 16, 16, 32 * nf), dev))
        o3 = self.dblocks[1].forward(self.dblocks[0].forward(cat3))
        cat2 = _S((b, 1, 32, 32, 20 * nf), dev)
        ops.resample_linear(o3, 4 * nf, cat2)
        l2.forward(feats["ra"][1], feats["re"][1], cat2)
        o2 = self.dblocks[3].forward(self.dblocks[2].forward(cat2))
        cat1 = _S((b, 1, 64, 64, 10 * nf), dev)
        ops.resample_linear(o2, 2 * nf, cat1)
        l1.forward(feats["ra"][0], feats["re"][0], cat1)
        sv.o1 = self.dblocks[5].forward(self.dblocks[4].forward(cat1))
        kpad = L.pad64(kp)
        logits = torch.empty((b, 64 * 64, kpad), dtype=torch.float32, device=dev)
        ops.conv_gemm(sv.o1, L.pad64(nf), self.head.w, kpad, out_f32=logits.view(b, 1, 64, 64, kpad), lcin=nf, lcout=kp)
        gb = self._gcn_buffers(b)
        rows = gb["rows"]
        heat = torch.empty((b, kp, 64, 64), dtype=torch.float32, device=dev)
        gcn = torch.empty_like(heat)
        st = gb["st"]                                                               # St_0, St_1, St_2 (inputs of the three layers)
        yt = [_S((1, 1, 1, rows, 1024), dev) for _ in range(2)]                     # Yt_1, Yt_2 (post-ReLU)
        for i in range(3):
            ops.gcn_bias_rows(self.gcn_b[i], b, gb["bias"][i])                      # bias[q][j] as residual rows [(b, j)][q]
        relu = self.gcn_relu
        ops.gcn_nodes(logits, self.adj, heat, st[0])
        ops.conv_gemm(st[0], 1024, self.gcn_w[0], 1024, residual=gb["bias"][0], slope=relu, out=yt[0], lrows=b * kp)
        ops.gcn_mix(yt[0], self.adj, st[1], b)
        ops.conv_gemm(st[1], 1024, self.gcn_w[1], 1024, residual=gb["bias"][1], slope=relu, out=yt[1], lrows=b * kp)
        ops.gcn_mix(yt[1], self.adj, st[2], b)
        y3 = torch.empty((1, 1, 1, rows, 1024), dtype=torch.float32, device=dev)
        ops.conv_gemm(st[2], 1024, self.gcn_w[2], 1024, residual=gb["bias"][2], out_f32=y3, lrows=b * kp)
        ops.gcn_heads(y3.view(rows, 1024), gcn, b)
        sv.st, sv.yt, sv.rows = st, yt, rows
        sv.heat, sv.gcn = heat, gcn
        self.last_outputs = (heat, gcn)
        return heat, gcn

    # ---------------------------------------------------------------------------------------------------------------- backward
    def _backward(self, dlogits, dpre):
        """dlogits float32 [B, 4096, kpad] (gradient of the head logits through the direct sigmoid path; the PRGCN path is ADDED to it),
        dpre float32 [B, 14, 64, 64] (gradient of the pre-sigmoid PRGCN maps) -> every parameter gradient, written into ``self.gview``."""
        sv = self._saved
        nf, g, kp = self.nf, self.g, self.kp
        dev = dlogits.device
        b, rows, st, yt = sv.b, sv.rows, sv.st, sv.yt
        kpad = L.pad64(kp)
        grads = dict(self.gview)
        grads[_UNPACK_JOBS], grads[_KEEP_ALIVE] = [], []
        gb = self._gcn_buffers(b)
        relu = self.gcn_relu
        dy3f = _zeros((rows, 1024), torch.float32, dev)
        ops._call("hupr_gcn_heads_bwd", dpre.data_ptr(), dy3f.data_ptr(), b, ops._C.stream_ptr())
        dyt = _S((1, 1, 1, rows, 1024), dev)
        T.accumulate(dyt, 1024, f=dy3f)
        p = "radarDecoder."
        for layer in (2, 1, 0):
            # dW[q][p] = sum_rows dYt[row][q] St[row][p]   (both operands transposed to rows-contiguous), written straight into the gradient
            a_t = ops.transpose_split(dyt, 1024, _S((1, 1024, rows), dev))
            s_t = ops.transpose_split(st[layer], 1024, _S((1, 1024, rows), dev))
            ops.conv_gemm(SplitTensor(a_t.hi.view(1, 1, 1, 1024, rows), a_t.lo.view(1, 1, 1, 1024, rows)), rows, s_t, 1024,
                          out_f32=grads[p + "gcn.L%d.weight" % (layer + 1)].view(1, 1, 1, 1024, 1024), lcin=b * kp)
            ops._call("hupr_gcn_bias_grad", dyt.hi.data_ptr(), dyt.lo.data_ptr(), grads[p + "gcn.L%d.bias" % (layer + 1)].data_ptr(), b,
                      ops._C.stream_ptr())
            dst = ops.conv_gemm(dyt, 1024, self.gcn_wt[layer], 1024, out=_S((1, 1, 1, rows, 1024), dev), lrows=b * kp)       # dSt = dYt W
            if layer == 0:
                ops._call("hupr_gcn_nodes_bwd", dst.hi.data_ptr(), dst.lo.data_ptr(), self.adj.data_ptr(), dlogits.data_ptr(), kpad, b, ops._C.stream_ptr())
            else:
                dy_prev = gb["dy"][layer - 1]
                ops.gcn_mix(dst, self.adj_t, dy_prev, b)                                                     # dY = dSt A^T (per sample)
                T.act_bwd(dy_prev, yt[layer - 1], 1024, relu, dy_prev)
                dyt = dy_prev
        # ---- head conv and decoder
        l3, l2, l1 = self.levels
        dlog = _S((b, 1, 64, 64, kpad), dev)
        T.accumulate(dlog, kpad, f=dlogits.view(-1, kpad))
        self.head.wgrad(sv.o1, 0, dlog, 0, grads)
        d = self.head.dgrad(dlog, _S((b, 1, 64, 64, L.pad64(nf)), dev))
        dcat1 = self.dblocks[4].backward(self.dblocks[5].backward(d, 0, grads), 0, grads)
        dra1, dre1 = l1.backward(dcat1, grads)
        acc = _zeros((b, 1, 32, 32, L.pad64(2 * nf)), torch.float32, dev)
        T.resample_linear_bwd((dcat1, 0), 2 * nf, acc)
        do2 = _S((b, 1, 32, 32, L.pad64(2 * nf)), dev)
        T.accumulate(do2, L.pad64(2 * nf), f=acc.view(-1, L.pad64(2 * nf)))
        dcat2 = self.dblocks[2].backward(self.dblocks[3].backward(do2, 0, grads), 0, grads)
        dra2, dre2 = l2.backward(dcat2, grads)
        acc = _zeros((b, 1, 16, 16, 4 * nf), torch.float32, dev)
        T.resample_linear_bwd((dcat2, 0), 4 * nf, acc)
        do3 = _S((b, 1, 16, 16, 4 * nf), dev)
        T.accumulate(do3, 4 * nf, f=acc.view(-1, 4 * nf))
        dcat3 = self.dblocks[0].backward(self.dblocks[1].backward(do3, 0, grads), 0, grads)
        dra3, dre3 = l3.backward(dcat3, grads)
        # ---- encoders and chirp nets
        for key, dfs, x in (("ra", (dra1, dra2, dra3), sv.hori), ("re", (dre1, dre2, dre3), sv.vert)):
            dchirp = self.enc[key].backward(dfs[0], dfs[1], dfs[2], grads)
            dfeat = _S((b, g, 64, 64, nf), dev)
            T.accumulate(dfeat, nf, a=(dchirp, 0))
            dwb = _zeros((nf * 4 + nf,), torch.float64, dev)
            ops._call("hupr_mnet_bwd", x.data_ptr(), self.mnet[key][0].data_ptr(), self.mnet[key][1].data_ptr(), dfeat.hi.data_ptr(),
                      dfeat.lo.data_ptr(), dwb.data_ptr(), dwb[nf * 4:].data_ptr(), b * g, ops._C.stream_ptr())
            net = "RAchirpNet" if key == "ra" else "REchirpNet"
            ops.reduce_f64(dwb, nf * 4, 1, grads[net + ".temporalConvWx1x1.weight"])
            ops.reduce_f64(dwb[nf * 4:], nf, 1, grads[net + ".temporalConvWx1x1.bias"])
        # ---- every filter gradient from its [taps, cin_pad, cout] accumulator into the torch-layout flat buffer: ONE launch.  The table is
        # rebuilt only when an accumulator moved (the planning pass of the scratch arena); it is static by the time a graph is captured.
        jobs = tuple(grads.pop(_UNPACK_JOBS))
        cached = self._unpack_tables.get(b)
        if cached is None or cached.key != jobs:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("hupr_b200.training: gradient accumulators moved during CUDA-graph capture (run the warm-up passes first)")
            cached = self._unpack_tables[b] = ops.PackTable(jobs, self.device)
        cached.unpack()
        grads.pop(_KEEP_ALIVE)
        self.last_grads = grads

    # ---------------------------------------------------------------------------------------------------------------- step
    def forward_backward(self, hori, vert, joints, loss_weights=(1.0, 1.0)):
        """One pass of /root/reference/tools/run.py:76-78 (model.train() forward, LossComputer loss, loss.backward()).  ``loss_weights`` =
        (alpha, beta) of loss = alpha*loss1 + beta*loss2 when TRAINING.lossDecay != -1 (misc/losses.py:36-42), (1, 1) otherwise.
        Returns the device scalars (loss1 + loss2, loss2)."""
        b = hori.shape[0]
        dev = hori.device
        kpad = L.pad64(self.kp)
        with self._scope(b):
            heat, gcn = self._forward(hori, vert)
            losses, _, _ = ops.heatmap_loss_fwd(heat, gcn, joints)
            dlogits = _zeros((b, 64 * 64, kpad), torch.float32, dev)
            dpre = torch.empty((b, self.kp, 64, 64), dtype=torch.float32, device=dev)
            ops.heatmap_loss_bwd(heat, gcn, joints, dlogits, dpre, weights=loss_weights)
            self._backward(dlogits, dpre)
        self.last_losses = losses
        return losses[0], losses[1]

    def forward(self, hori, vert):
        """Train-mode forward only (the autograd bridge of HuPRNet.forward): (heatmap, gcn_heatmap) float32 [B,14,64,64]."""
        with self._scope(hori.shape[0], arena=False):
            return self._forward(hori, vert)

    def backward(self, g_heat, g_gcn):
        """Backward of the last ``forward`` for caller-supplied output gradients dL/d heatmap, dL/d gcn_heatmap ([B,14,64,64] or None)."""
        sv = self._saved
        b, dev = sv.b, sv.heat.device
        kpad = L.pad64(self.kp)
        with self._scope(b, arena=False):
            dlogits = _zeros((b, 64 * 64, kpad), torch.float32, dev)
            dpre = torch.empty((b, self.kp, 64, 64), dtype=torch.float32, device=dev)
            gh = None if g_heat is None else g_heat.reshape(b, self.kp, 64, 64).float().contiguous()
            gg = None if g_gcn is None else g_gcn.reshape(b, self.kp, 64, 64).float().contiguous()
            ops.heatmap_bwd(sv.heat, sv.gcn, gh, gg, dlogits, dpre)
            self._backward(dlogits, dpre)

    def optimizer_step(self):
        self.step_count += 1
        h = self.hyper
        if not torch.cuda.is_current_stream_capturing():
            self._sync_lr()
        ops.bump_i32(self.step_dev)        # device-side step counter: the bias corrections stay right under CUDA-graph replay
        ops.adam_step(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, self.step_count, lr=h["lr"], betas=h["betas"], eps=h["eps"],
                      weight_decay=h["weight_decay"], step_dev=self.step_dev, lr_dev=self.lr_dev)
        self._dirty = True
        self.model.invalidate()

    def capture(self, hori, vert, joints, with_optimizer=True, loss_weights=(1.0, 1.0)):
        """Capture forward_backward (+ the Adam launch) into one CUDA graph over the caller's STATIC input tensors; returns a
        ``replay()`` callable whose result is the device tensor pair (loss, loss2).  The step is ~1 000 launches, so at small
        batches the Python/ctypes launch path is the bottleneck; the graph removes it.  The learning rate and the Adam step count are
        read from device scalars, so schedules keep working across replays.  (With several ranks capture only
        ``with_optimizer=False`` and run all_reduce_gradients() + optimizer_step() eagerly after each replay.)"""
        joints = joints.to(device=hori.device, dtype=torch.int64).contiguous()
        for _ in range(2):                       # warm-up: lazy buffers, the scratch arena's planning pass, allocator high-water mark
            self.forward_backward(hori, vert, joints, loss_weights)
            if with_optimizer:
                self.optimizer_step()
        torch.cuda.synchronize()
        self._dirty = True                       # the operand re-pack is part of the graph: replays see the weights of the latest Adam launch
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.forward_backward(hori, vert, joints, loss_weights)
            if with_optimizer:
                self.optimizer_step()
        self._graph = graph
        if with_optimizer:
            self.step_count -= 1                 # the captured optimizer_step() call was recorded, not executed

        def replay():
            if with_optimizer:
                self.step_count += 1             # the device-side counter (step_dev) is advanced by the captured graph itself
                self._sync_lr()
            graph.replay()
            return out
        return replay

    def all_reduce_gradients(self):
        """Data-parallel training (SURVEY.md §8 e): ONE sum all-reduce over the flat gradient buffer, then divide by the world size."""
        from .sharding import average_gradients
        average_gradients(self.flat_g)

    # ---------------------------------------------------------------------------------------------------------------- ingest
    def prefetch(self, hori, vert, joints):
        """Start the asynchronous upload of the NEXT batch (pinned host tensors) on a copy stream while the current step computes — the
        training-side counterpart of RadarPoseStream.prefetch.  ``step_prefetched()`` then trains on it."""
        dev = self.device
        st = getattr(self, "_ingest", None)
        if st is None or st["hori"].shape != hori.shape:
            st = self._ingest = dict(hori=torch.empty(hori.shape, dtype=torch.float32, device=dev),
                                     vert=torch.empty(vert.shape, dtype=torch.float32, device=dev),
                                     joints=torch.empty(joints.shape, dtype=torch.int64, device=dev),
                                     stream=torch.cuda.Stream(device=dev), uploaded=torch.cuda.Event(), consumed=torch.cuda.Event())
            st["consumed"].record(torch.cuda.current_stream(dev))
        st["stream"].wait_event(st["consumed"])          # the previous staging contents have been handed to the step
        with torch.cuda.stream(st["stream"]):
            st["hori"].copy_(hori, non_blocking=True)
            st["vert"].copy_(vert, non_blocking=True)
            st["joints"].copy_(joints, non_blocking=True)
            st["uploaded"].record(st["stream"])

    def prefetch_adc(self, adc_hori, adc_vert, joints):
        """Compact ingest: upload the NEXT batch as raw DCA1000 words — int16 ``[B, 8, FRAME_WORDS]`` per sensor (the eight frames of each
        sample's window; pinned host tensors), 12.6 MB per sample instead of the 32 MiB of float32 VRDAE maps — on the copy stream.
        ``take_prefetched_adc`` then runs the FFT cascade and the window standardisation on the device (what
        preprocessing/process_iwr1843.py:106-173 and datasets/dataset.py:139-150 do on the host in the reference)."""
        from .preprocessing.process_iwr1843 import FRAME_WORDS
        dev = self.device
        b, g = adc_hori.shape[0], adc_hori.shape[1]
        st = getattr(self, "_ingest_adc", None)
        if st is None or st["adc"].shape[1] != b * g:
            st = self._ingest_adc = dict(adc=torch.empty((2, b * g, FRAME_WORDS), dtype=torch.int16, device=dev),
                                         cubes=torch.empty((2 * b * g, 16, 64, 64, 8), dtype=torch.complex64, device=dev),
                                         slots=[torch.arange(b * g, dtype=torch.int32, device=dev),
                                                torch.arange(b * g, 2 * b * g, dtype=torch.int32, device=dev)],
                                         joints=torch.empty(joints.shape, dtype=torch.int64, device=dev),
                                         stream=torch.cuda.Stream(device=dev), uploaded=torch.cuda.Event(), consumed=torch.cuda.Event())
            st["consumed"].record(torch.cuda.current_stream(dev))
        st["stream"].wait_event(st["consumed"])
        with torch.cuda.stream(st["stream"]):
            st["adc"][0].copy_(adc_hori.view(b * g, FRAME_WORDS), non_blocking=True)
            st["adc"][1].copy_(adc_vert.view(b * g, FRAME_WORDS), non_blocking=True)
            st["joints"].copy_(joints, non_blocking=True)
            st["uploaded"].record(st["stream"])

    def take_prefetched_adc(self, hori, vert, joints):
        """FFT cascade + window standardisation of the batch uploaded by ``prefetch_adc`` into the step's static VRDAE input tensors."""
        from .preprocessing.process_iwr1843 import FRAME_WORDS, cascade_i16
        st = self._ingest_adc
        main = torch.cuda.current_stream(self.device)
        main.wait_event(st["uploaded"])
        n = st["adc"].shape[1]
        cascade_i16(st["adc"].view(2 * n, FRAME_WORDS), st["cubes"])
        joints.copy_(st["joints"], non_blocking=True)
        st["consumed"].record(main)                       # the staging words are consumed; the cubes are this stream's own scratch
        ops.window_normalize(st["cubes"], st["slots"][0], hori.view(n, 8, 2, 64, 64, 8))
        ops.window_normalize(st["cubes"], st["slots"][1], vert.view(n, 8, 2, 64, 64, 8))

    def take_prefetched(self, hori, vert, joints):
        """Move the batch uploaded by the last ``prefetch`` into the step's static input tensors (device-to-device, stream-ordered)."""
        st = self._ingest
        main = torch.cuda.current_stream(self.device)
        main.wait_event(st["uploaded"])
        hori.copy_(st["hori"], non_blocking=True)
        vert.copy_(st["vert"], non_blocking=True)
        joints.copy_(st["joints"], non_blocking=True)
        st["consumed"].record(main)
