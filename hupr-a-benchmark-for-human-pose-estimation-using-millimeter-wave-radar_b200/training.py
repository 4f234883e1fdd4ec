"""Training step of MSCSA-PRGCN on the B200: train-mode forward (BatchNorm batch statistics), hand-assembled
backward pass and Adam, every arithmetic step a launch into libhupr_b200.so.

Replaces ``loss.backward(); optimizer.step()`` of /root/reference/tools/run.py:76-79 with ``optim.Adam(lr 1e-4, betas (.9,.999),
weight_decay 1e-4)`` of tools/base.py:47.  There is no autograd graph: the backward of each stage is written out against the
tensors the forward saved —
  convolutions : dgrad = the implicit-GEMM kernel with flipped filters; wgrad = hupr_conv_wgrad: positions contracted on the tensor
                 cores with MN-major operands read straight from the channels-last tensors — ops.conv_wgrad_direct
  BatchNorm    : batch statistics (double sums), affine + ReLU, standard three-term backward — train_ops
  attention    : P is recomputed (QK^T GEMM + row softmax), then dV = P^T dO, dP = dO V^T, dS = P*(dP - rowsum), dQ = dS K, dK = dS^T Q
                 (the two A^T B products contract over queries with MN-major operands — ops.matmul_tn, no transposed [S,S] copies)
  PRGCN        : the transposed-layout GEMMs of the forward with W^T, adjacency mixes with A^T, adjoint resampling
This first version favours exactness over speed (fp32-equivalent hi/lo arithmetic everywhere, unfused backward attention).
"""
import os

import torch

from . import ops
from . import train_ops as T
from .models import layers as L
from .models.networks import ADJACENCY
from .ops import SplitTensor

EPS = 1e-5
MOMENTUM = 0.1


def _S(shape, dev, zero=False):
    return SplitTensor.empty(tuple(shape), dev, True, zero=zero)


def _rows(t):
    """SplitTensor [n, d, h, w, c] -> view [n, 1, 1, d*h*w, c]."""
    n, c = t.hi.shape[0], t.hi.shape[-1]
    s = t.hi.numel() // (n * c)
    return SplitTensor(t.hi.view(n, 1, 1, s, c), t.lo.view(n, 1, 1, s, c))


class ConvOp(object):
    """One (possibly cout-concatenated) stride-1 convolution: forward / dgrad / wgrad packs and gradient bookkeeping."""

    def __init__(self, names, cin, couts, kernel, pad, bias_name=None):
        self.names, self.cin, self.couts = names, cin, couts
        self.kernel, self.pad, self.bias_name = tuple(kernel), tuple(pad), bias_name
        self.cin_pad = L.pad64(cin)
        self.cout_pads = [L.pad64(c) for c in couts]
        self.cout = sum(self.cout_pads)
        self.taps = kernel[0] * kernel[1] * kernel[2]

    def _w5(self, w):
        return w if w.dim() == 5 else w.unsqueeze(2)

    def pack(self, sd):
        ws = [self._w5(sd[n].detach().float()) for n in self.names]
        self.w = L.pack_conv(ws, self.cin_pad, self.cout_pads)
        full = torch.zeros((self.cout, self.cin_pad) + self.kernel, dtype=torch.float32, device=ws[0].device)
        o = 0
        for w, cp in zip(ws, self.cout_pads):
            full[o:o + w.shape[0], :w.shape[1]] = w
            o += cp
        self.wd = L.pack_dgrad(full, self.cout, self.cin_pad)
        self.bias = sd[self.bias_name].detach().float().contiguous() if self.bias_name else None

    def forward(self, x, out, **kw):
        return ops.conv_gemm(x, self.cin_pad, self.w, self.cout, kernel=self.kernel, pad=self.pad, shift=self.bias, out=out, **kw)

    def dgrad(self, dy, out, **kw):
        """dy: SplitTensor [..., self.cout] (or a wider tensor with a_ch_off) -> out [..., cin_pad]."""
        dpad = (self.kernel[0] - 1 - self.pad[0], self.pad[1], self.pad[2])
        return ops.conv_gemm(dy, self.cout, self.wd, self.cin_pad, kernel=self.kernel, pad=dpad, out=out, **kw)

    def wgrad(self, x, x_off, dy, dy_off, grads):
        """x: saved input [n, d, h, w, ld]; dy: output gradient [n, d_out, h, w, ld'] -> grads[name] (+ bias gradient).
        One hupr_conv_wgrad launch: positions are contracted on the tensor cores straight from the channels-last tensors."""
        dev = x.hi.device
        acc = ops.conv_wgrad_direct(x, x_off, self.cin_pad, dy, dy_off, self.cout, self.kernel, self.pad)      # [taps, cin_pad, cout]
        o = 0
        for name, c, cp in zip(self.names, self.couts, self.cout_pads):
            grads[name] = acc[:, :self.cin, o:o + c].permute(2, 1, 0)      # strided view [c, cin, taps]: copied once, into the flat gradient buffer
            o += cp
        if self.bias_name:
            s1 = torch.zeros(self.cout, dtype=torch.float64, device=dev)
            s2 = torch.zeros_like(s1)
            T.channel_sums(T.SUMS_PRELU, (dy, dy_off), self.cout, s1, s2)
            grads[self.bias_name] = s2[:self.couts[0]].float()


class BNOp(object):
    def __init__(self, prefix, c):
        self.prefix, self.c = prefix, c

    def forward_stats(self, z, z_off, params, buffers, count):
        """Batch statistics of z[..., z_off:z_off+c]; updates running stats; returns (scale, shift) for the affine kernel."""
        dev = z.hi.device
        sums = torch.zeros((2, self.c), dtype=torch.float64, device=dev)
        T.channel_sums(T.SUMS_STATS, (z, z_off), self.c, sums[0], sums[1])
        gamma, beta = params[self.prefix + ".weight"].detach(), params[self.prefix + ".bias"].detach()
        rm, rv = buffers[self.prefix + ".running_mean"], buffers[self.prefix + ".running_var"]
        nbt = buffers[self.prefix + ".num_batches_tracked"]
        ok = all(t.dtype == torch.float32 and t.is_contiguous() for t in (gamma, beta, rm, rv)) and nbt.dtype == torch.int64
        if ok:      # one launch: mean, rstd, fused affine, running statistics (hupr_bn_finalize)
            st = T.bn_finalize(sums, count, gamma, beta, EPS, MOMENTUM, rm, rv, nbt)
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
        sums = torch.zeros((2, self.c), dtype=torch.float64, device=dev)
        T.channel_sums(T.SUMS_BN_BWD, g, self.c, sums[0], sums[1], b=(z, z_off), mask=mask, mean=self.mean, rstd=self.rstd)
        st = T.bn_bwd_finalize(sums, count)                      # k2, k3, dgamma, dbeta in one launch
        T.bn_bwd_apply(g, (z, z_off), self.c, self.mean, self.rstd, self.scale, st[0], st[1], (out, out_off), mask=mask)
        grads[self.prefix + ".weight"] = st[2]
        grads[self.prefix + ".bias"] = st[3]


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
        self.zc = self.conv1.forward(x, _S(shape + (2 * c,), dev))
        s1, h1 = self.bn1.forward_stats(self.zc, 0, params, buffers, count)
        sd_, hd = self.bnd.forward_stats(self.zc, c, params, buffers, count)
        self.t = _S(shape + (c,), dev)
        T.affine_act((self.zc, 0), c, self.t, scale1=s1, shift1=h1, slope=self.relu)
        self.z2 = self.conv2.forward(self.t, _S(shape + (c,), dev))
        s2, h2 = self.bn2.forward_stats(self.z2, 0, params, buffers, count)
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

    def pack(self, sd):
        self.conv1.pack(sd)
        self.conv2.pack(sd)
        dev = sd[self.prefix + ".relu.weight"].device
        self.a1 = sd[self.prefix + ".main.1.weight"].detach().float().expand(self.cp).contiguous()
        self.a2 = sd[self.prefix + ".relu.weight"].detach().float().expand(self.cp).contiguous()

    def forward(self, x, out=None, out_off=0):
        dev, cp = x.hi.device, self.cp
        shape = x.hi.shape[:-1]
        self.x = x
        self.zc = self.conv1.forward(x, _S(shape + (2 * cp,), dev))
        self.t = _S(shape + (cp,), dev)
        T.affine_act((self.zc, 0), cp, self.t, slope=self.a1)
        self.s = ops.conv_gemm(self.t, cp, self.conv2.w, cp, kernel=(1, 3, 3), pad=(0, 1, 1), residual=self.zc, r_ch_off=cp, out=_S(shape + (cp,), dev))
        self.out = _S(shape + (cp,), dev) if out is None else out
        T.affine_act(self.s, cp, (self.out, out_off), slope=self.a2)
        return self.out

    def backward(self, dout, dout_off, grads):
        dev, cp = self.s.hi.device, self.cp
        shape = self.s.hi.shape[:-1]
        ds = _S(shape + (cp,), dev)
        T.act_bwd((dout, dout_off), self.s, cp, self.a2, ds)
        p1 = torch.zeros(cp, dtype=torch.float64, device=dev)
        p2 = torch.zeros_like(p1)
        T.channel_sums(T.SUMS_PRELU, (dout, dout_off), cp, p1, p2, b=self.s)
        grads[self.prefix + ".relu.weight"] = p1.sum().float().reshape(1)
        self.conv2.wgrad(self.t, 0, ds, 0, grads)
        dt = self.conv2.dgrad(ds, _S(shape + (cp,), dev))
        dzc = _S(shape + (2 * cp,), dev)
        T.act_bwd(dt, (self.zc, 0), cp, self.a1, (dzc, 0))
        q1 = torch.zeros(cp, dtype=torch.float64, device=dev)
        q2 = torch.zeros_like(q1)
        T.channel_sums(T.SUMS_PRELU, dt, cp, q1, q2, b=(self.zc, 0))
        grads[self.prefix + ".main.1.weight"] = q1.sum().float().reshape(1)
        T.accumulate((dzc, cp), cp, a=ds)
        self.conv1.wgrad(self.x, 0, dzc, 0, grads)
        return self.conv1.dgrad(dzc, _S(shape + (self.conv1.cin_pad,), dev))


class AttentionLevel(object):
    """One scale of the cross/self attention (layers.py:126-149) with its eight 1x1 projections."""

    def __init__(self, level, c, hw, prev):
        p = "radarDecoder."
        self.c, self.hw, self.s, self.prev = c, hw, hw * hw, prev
        self.fused_bwd = False       # hupr_attention_bwd for head dim 64 (opt-in until it has been timed inside the step: DESIGN.md §3b)
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
        dvf = {"ra": torch.zeros((b, s, c), dtype=torch.float32, device=dev), "re": torch.zeros((b, s, c), dtype=torch.float32, device=dev)}
        dv_res = {}
        if self.fused_bwd and c == 64 and all(l is not None for l in self.lse):
            # fused flash-style backward: no [S, S] matrix in memory; dQ / dK of the four attentions accumulate in fp32 projection-gradient
            # buffers (one conversion to hi/lo afterwards), dV in the per-source buffers
            dprf = {key: torch.zeros((b, s, 4 * c), dtype=torch.float32, device=dev) for key in ("ra", "re")}
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
            dkf = torch.zeros((b, s, c), dtype=torch.float32, device=dev)
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
        acc = torch.zeros((b, g // 2, 32, 32, 4 * nf), dtype=torch.float32, device=dev)
        T.resample_linear_bwd(d, 4 * nf, acc)
        self.merges[1].wgrad(self.l2, 0, df2, 0, grads)
        dl2 = self.merges[1].dgrad(df2, _S((b, g // 2, 32, 32, 4 * nf), dev))
        T.accumulate(dl2, 4 * nf, a=dl2, f=acc.view(-1, 4 * nf))
        d = self.blocks[1].backward(self.blocks[2].backward(dl2, grads), grads)                 # -> d l2in [B, 4, 32, 32, 64]
        acc = torch.zeros((b, g, 64, 64, 2 * nf), dtype=torch.float32, device=dev)
        T.resample_linear_bwd(d, 2 * nf, acc)
        self.merges[0].wgrad(self.l1, 0, df1, 0, grads)
        dl1 = self.merges[0].dgrad(df1, _S((b, g, 64, 64, 2 * nf), dev))
        T.accumulate(dl1, 2 * nf, a=dl1, f=acc.view(-1, 2 * nf))
        dl1a = self.blocks[0].backward(dl1, grads)
        self.conv0.wgrad(self.x0, 0, dl1a, 0, grads)
        return self.conv0.dgrad(dl1a, _S((b, g, 64, 64, L.pad64(nf)), dev))


class TrainStep(object):
    """``loss, loss2 = step.forward_backward(hori, vert, joints)`` fills ``param.grad`` of every model parameter;
    ``step.optimizer_step()`` applies Adam (coupled L2) in place.  ``model`` is a ``hupr_b200.models.HuPRNet`` on a CUDA device."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, fused_attention_bwd=None):
        """``fused_attention_bwd``: use hupr_attention_bwd for the head-dim-64 attentions (default: off unless HUPR_FUSED_ATTN_BWD=1 —
        the kernel is parity-tested on its own but has not been run inside the step yet)."""
        self.model = model
        nf, g, kp = model.numFilters, model.numGroupFrames, model.numKeypoints
        self.nf, self.g, self.kp = nf, g, kp
        self.enc = {"ra": Encoder("RAradarEncoder", nf, g), "re": Encoder("REradarEncoder", nf, g)}
        self.levels = [AttentionLevel(0, 8 * nf, 16, 0), AttentionLevel(1, 4 * nf, 32, 4 * nf), AttentionLevel(2, 2 * nf, 64, 2 * nf)]
        if fused_attention_bwd is None:
            fused_attention_bwd = os.environ.get("HUPR_FUSED_ATTN_BWD") == "1"
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
        n = sum(q.numel() for q in params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        for q in params:
            k = q.numel()
            self.flat_p[o:o + k].copy_(q.data.reshape(-1))
            q.data = self.flat_p[o:o + k].view(q.shape)
            q.grad = self.flat_g[o:o + k].view(q.shape)
            o += k
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.adj = torch.tensor(ADJACENCY, dtype=torch.float32, device=dev)
        self.adj_t = self.adj.t().contiguous()
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._dirty = True

    # ---------------------------------------------------------------------------------------------------------------- packing
    def _pack(self):
        sd = {k: v for k, v in self.model.named_parameters()}
        for e in self.enc.values():
            e.pack(sd)
        for m in self.levels + self.dblocks:
            m.pack(sd)
        self.head.pack(sd)
        p = "radarDecoder."
        self.gcn_w = [SplitTensor.from_float(sd[p + "gcn.L%d.weight" % i].detach().float().view(1, 1024, 1024)) for i in (1, 2, 3)]
        self.gcn_wt = [SplitTensor.from_float(sd[p + "gcn.L%d.weight" % i].detach().float().t().contiguous().view(1, 1024, 1024)) for i in (1, 2, 3)]
        self.gcn_b = [sd[p + "gcn.L%d.bias" % i].detach().float() for i in (1, 2, 3)]
        self.mnet = {k: (sd[n + ".temporalConvWx1x1.weight"].detach().float().contiguous(), sd[n + ".temporalConvWx1x1.bias"].detach().float().contiguous())
                     for k, n in (("ra", "RAchirpNet"), ("re", "REchirpNet"))}
        self._dirty = False

    # ---------------------------------------------------------------------------------------------------------------- step
    def forward_backward(self, hori, vert, joints):
        if self._dirty:
            self._pack()
        model, nf, g, kp = self.model, self.nf, self.g, self.kp
        dev = hori.device
        b = hori.shape[0]
        params = dict(model.named_parameters())
        buffers = dict(model.named_buffers())
        grads = {}
        # ---- forward
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
        o1 = self.dblocks[5].forward(self.dblocks[4].forward(cat1))
        kpad = L.pad64(kp)
        logits = torch.empty((b, 64 * 64, kpad), dtype=torch.float32, device=dev)
        ops.conv_gemm(o1, L.pad64(nf), self.head.w, kpad, out_f32=logits.view(b, 1, 64, 64, kpad))
        rows = -(-(b * kp) // 128) * 128
        heat = torch.empty((b, kp, 64, 64), dtype=torch.float32, device=dev)
        gcn = torch.empty_like(heat)
        st = [_S((1, 1, 1, rows, 1024), dev, zero=True) for _ in range(3)]          # St_0, St_1, St_2 (inputs of the three layers)
        yt = [_S((1, 1, 1, rows, 1024), dev, zero=True) for _ in range(2)]          # Yt_1, Yt_2 (post-ReLU)
        bias_rows = []
        for bias in self.gcn_b:
            t = torch.zeros(rows, 1024, dtype=torch.float32, device=dev)
            t[:b * kp] = bias.t().repeat(b, 1)
            bias_rows.append(SplitTensor.from_float(t.view(1, 1, 1, rows, 1024)))
        relu = torch.zeros(1024, dtype=torch.float32, device=dev)
        ops.gcn_nodes(logits, self.adj, heat, st[0])
        ops.conv_gemm(st[0], 1024, self.gcn_w[0], 1024, residual=bias_rows[0], slope=relu, out=yt[0])
        ops.gcn_mix(yt[0], self.adj, st[1], b)
        ops.conv_gemm(st[1], 1024, self.gcn_w[1], 1024, residual=bias_rows[1], slope=relu, out=yt[1])
        ops.gcn_mix(yt[1], self.adj, st[2], b)
        y3 = torch.zeros((1, 1, 1, rows, 1024), dtype=torch.float32, device=dev)
        ops.conv_gemm(st[2], 1024, self.gcn_w[2], 1024, residual=bias_rows[2], out_f32=y3)
        ops.gcn_heads(y3.view(rows, 1024), gcn, b)
        losses, _, _ = ops.heatmap_loss_fwd(heat, gcn, joints)
        self.last_outputs = (heat, gcn)
        # ---- backward: loss -> logits (two paths: direct sigmoid head, PRGCN)
        dlogits = torch.zeros((b, 64 * 64, kpad), dtype=torch.float32, device=dev)
        dpre = torch.empty((b, kp, 64, 64), dtype=torch.float32, device=dev)
        ops.heatmap_loss_bwd(heat, gcn, joints, dlogits, dpre)
        dy3f = torch.zeros((rows, 1024), dtype=torch.float32, device=dev)
        ops._call("hupr_gcn_heads_bwd", dpre.data_ptr(), dy3f.data_ptr(), b, ops._C.stream_ptr())
        dyt = _S((1, 1, 1, rows, 1024), dev)
        T.accumulate(dyt, 1024, f=dy3f)
        p = "radarDecoder."
        for layer in (2, 1, 0):
            # dW[q][p] = sum_rows dYt[row][q] St[row][p]   (both operands transposed to rows-contiguous)
            a_t = ops.transpose_split(dyt, 1024, _S((1, 1024, rows), dev))
            s_t = ops.transpose_split(st[layer], 1024, _S((1, 1024, rows), dev))
            dw = torch.empty((1, 1, 1, 1024, 1024), dtype=torch.float32, device=dev)
            ops.conv_gemm(SplitTensor(a_t.hi.view(1, 1, 1, 1024, rows), a_t.lo.view(1, 1, 1, 1024, rows)), rows, s_t, 1024, out_f32=dw)
            grads[p + "gcn.L%d.weight" % (layer + 1)] = dw.view(1024, 1024)
            db = torch.empty((1024, kp), dtype=torch.float32, device=dev)
            ops._call("hupr_gcn_bias_grad", dyt.hi.data_ptr(), dyt.lo.data_ptr(), db.data_ptr(), b, ops._C.stream_ptr())
            grads[p + "gcn.L%d.bias" % (layer + 1)] = db
            dst = ops.conv_gemm(dyt, 1024, self.gcn_wt[layer], 1024, out=_S((1, 1, 1, rows, 1024), dev))       # dSt = dYt W
            if layer == 0:
                ops._call("hupr_gcn_nodes_bwd", dst.hi.data_ptr(), dst.lo.data_ptr(), self.adj.data_ptr(), dlogits.data_ptr(), kpad, b, ops._C.stream_ptr())
            else:
                dy_prev = _S((1, 1, 1, rows, 1024), dev, zero=True)
                ops.gcn_mix(dst, self.adj_t, dy_prev, b)                                                     # dY = dSt A^T (per sample)
                T.act_bwd(dy_prev, yt[layer - 1], 1024, relu, dy_prev)
                dyt = dy_prev
        # ---- head conv and decoder
        dlog = _S((b, 1, 64, 64, kpad), dev)
        T.accumulate(dlog, kpad, f=dlogits.view(-1, kpad))
        self.head.wgrad(o1, 0, dlog, 0, grads)
        d = self.head.dgrad(dlog, _S((b, 1, 64, 64, L.pad64(nf)), dev))
        dcat1 = self.dblocks[4].backward(self.dblocks[5].backward(d, 0, grads), 0, grads)
        dra1, dre1 = l1.backward(dcat1, grads)
        acc = torch.zeros((b, 1, 32, 32, L.pad64(2 * nf)), dtype=torch.float32, device=dev)
        T.resample_linear_bwd((dcat1, 0), 2 * nf, acc)
        do2 = _S((b, 1, 32, 32, L.pad64(2 * nf)), dev)
        T.accumulate(do2, L.pad64(2 * nf), f=acc.view(-1, L.pad64(2 * nf)))
        dcat2 = self.dblocks[2].backward(self.dblocks[3].backward(do2, 0, grads), 0, grads)
        dra2, dre2 = l2.backward(dcat2, grads)
        acc = torch.zeros((b, 1, 16, 16, 4 * nf), dtype=torch.float32, device=dev)
        T.resample_linear_bwd((dcat2, 0), 4 * nf, acc)
        do3 = _S((b, 1, 16, 16, 4 * nf), dev)
        T.accumulate(do3, 4 * nf, f=acc.view(-1, 4 * nf))
        dcat3 = self.dblocks[0].backward(self.dblocks[1].backward(do3, 0, grads), 0, grads)
        dra3, dre3 = l3.backward(dcat3, grads)
        # ---- encoders and chirp nets
        for key, dfs, x in (("ra", (dra1, dra2, dra3), hori), ("re", (dre1, dre2, dre3), vert)):
            dchirp = self.enc[key].backward(dfs[0], dfs[1], dfs[2], grads)
            dfeat = _S((b, g, 64, 64, nf), dev)
            T.accumulate(dfeat, nf, a=(dchirp, 0))
            dw = torch.zeros(nf * 4, dtype=torch.float64, device=dev)
            db = torch.zeros(nf, dtype=torch.float64, device=dev)
            ops._call("hupr_mnet_bwd", x.data_ptr(), self.mnet[key][0].data_ptr(), self.mnet[key][1].data_ptr(), dfeat.hi.data_ptr(),
                      dfeat.lo.data_ptr(), dw.data_ptr(), db.data_ptr(), b * g, ops._C.stream_ptr())
            net = "RAchirpNet" if key == "ra" else "REchirpNet"
            grads[net + ".temporalConvWx1x1.weight"] = dw.float().view(nf, 2, 2, 1, 1)
            grads[net + ".temporalConvWx1x1.bias"] = db.float()
        # ---- scatter into the flat gradient buffer (views installed as param.grad)
        for name, q in params.items():
            g = grads[name]
            if g.shape != q.shape and g.numel() == q.numel() and not g.is_contiguous():
                q.grad.view(g.shape).copy_(g)                     # e.g. a conv filter gradient view [c, cin, taps] -> [c, cin, kd, kh, kw]
            else:
                q.grad.copy_(g.reshape(q.shape))
        self.last_grads = grads
        return losses[0], losses[1]

    def optimizer_step(self):
        self.step_count += 1
        h = self.hyper
        self.step_dev.add_(1)          # device-side step counter: the bias corrections stay right under CUDA-graph replay
        ops.adam_step(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, self.step_count, lr=h["lr"], betas=h["betas"], eps=h["eps"],
                      weight_decay=h["weight_decay"], step_dev=self.step_dev)
        self._dirty = True
        self.model.invalidate()

    def capture(self, hori, vert, joints, with_optimizer=True):
        """Capture forward_backward (+ the Adam launch) into one CUDA graph over the caller's STATIC input tensors; returns a
        ``replay()`` callable whose result is the device tensor pair (loss, loss2).  The step launches ~1 150 kernels, so at small
        batches the Python/ctypes launch path is the bottleneck; the graph removes it.  (With several ranks capture only
        ``with_optimizer=False`` and run all_reduce_gradients() + optimizer_step() eagerly after each replay.)"""
        joints = joints.to(device=hori.device, dtype=torch.int64).contiguous()
        for _ in range(2):                       # warm-up: lazy function attributes, allocator high-water mark
            self.forward_backward(hori, vert, joints)
            if with_optimizer:
                self.optimizer_step()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.forward_backward(hori, vert, joints)
            if with_optimizer:
                self.optimizer_step()
        self._graph = graph
        if with_optimizer:
            self.step_count -= 1                 # the captured optimizer_step() call was recorded, not executed

        def replay():
            if with_optimizer:
                self.step_count += 1             # the device-side counter (step_dev) is advanced by the captured graph itself
            graph.replay()
            return out
        return replay

    def all_reduce_gradients(self):
        """Data-parallel training (SURVEY.md §8 e): ONE sum all-reduce over the flat gradient buffer, then divide by the world size."""
        from .sharding import average_gradients
        average_gradients(self.flat_g)
