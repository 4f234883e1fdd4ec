"""Weight packing and the launch sequences of the encoder / decoder stages (the reference's models/layers.py,
models/chirp_networks.py and models/gcn_networks.py, restated as kernel launches).

Layout conventions:
  * activations: channels-last ``[B, D, H, W, C]`` bf16 hi/lo split tensors (``ops.SplitTensor``), value = hi + lo;
  * conv weights: ``[taps, cout, cin]`` split tensors, cin/cout zero-padded to multiples of 64;
  * eval-mode BatchNorm is applied as an fp32 per-channel scale/shift in the conv epilogue (weights are not rescaled);
  * the two convolutions that read the same input in a residual block (``main.0`` and ``downsample.0``) are one launch
    with their filters concatenated along cout; a per-channel slope array gives ReLU/PReLU to the first half and identity to
    the second half, and the block's second conv fuses ``+ residual`` and the final activation into its epilogue.
"""
import torch

from .. import ops
from ..ops import SplitTensor

BN_EPS = 1e-5


def pad64(c):
    return -(-c // 64) * 64


def pack_conv(weights, cin_pad, cout_pads, split=True):
    """List of torch conv weights ``[Cout, Cin, *k]`` (same Cin, same kernel) -> SplitTensor ``[taps, sum(cout_pads), cin_pad]``."""
    mats = []
    for w, cp in zip(weights, cout_pads):
        cout, cin = w.shape[:2]
        taps = w[0, 0].numel()
        m = torch.zeros(taps, cp, cin_pad, dtype=torch.float32, device=w.device)
        m[:, :cout, :cin] = w.detach().float().reshape(cout, cin, taps).permute(2, 0, 1)
        mats.append(m)
    return SplitTensor.from_float(torch.cat(mats, dim=1), split)


def pack_dgrad(weight, cout_pad, cin_pad, split=True):
    """Data-gradient filter of a stride-1 'same' convolution: ``dX = conv(dY, W')`` with ``W'[tap][ci][co] = W[co][ci][flipped tap]``
    — the same implicit-GEMM kernel computes dgrad (reference: autograd of nn.Conv3d / nn.Conv2d).  For a depth-valid kernel
    ``(T, 1, 1)`` call the convolution with ``pad=(T-1, 0, 0)``."""
    flip_dims = tuple(range(2, weight.dim()))
    return pack_conv([weight.transpose(0, 1).flip(flip_dims)], cout_pad, [cin_pad], split)


def bn_affine(sd, prefix):
    scale = sd[prefix + ".weight"].float() / torch.sqrt(sd[prefix + ".running_var"].float() + BN_EPS)
    shift = sd[prefix + ".bias"].float() - sd[prefix + ".running_mean"].float() * scale
    return scale, shift


class Block(object):
    """Packed residual block: conv1 || downsample, then conv2 (+ residual, activation)."""
    __slots__ = ("w1", "scale1", "shift1", "slope1", "w2", "scale2", "shift2", "slope2", "cin", "cmid", "taps", "pad", "lc")      # lc: un-padded width


def pack_block3d(sd, prefix, cin, cout, split):
    b = Block()
    b.cin, b.cmid, b.taps, b.pad, b.lc = pad64(cin), cout, (3, 3, 3), (1, 1, 1), cout
    b.w1 = pack_conv([sd[prefix + ".main.0.weight"], sd[prefix + ".downsample.0.weight"]], b.cin, [cout, cout], split)
    s1, h1 = bn_affine(sd, prefix + ".main.1")
    sd_, hd = bn_affine(sd, prefix + ".downsample.1")
    b.scale1, b.shift1 = torch.cat([s1, sd_]).contiguous(), torch.cat([h1, hd]).contiguous()
    b.slope1 = torch.cat([torch.zeros_like(s1), torch.ones_like(sd_)]).contiguous()      # ReLU | identity
    b.w2 = pack_conv([sd[prefix + ".main.3.weight"]], cout, [cout], split)
    b.scale2, b.shift2 = (t.contiguous() for t in bn_affine(sd, prefix + ".main.4"))
    b.slope2 = torch.zeros_like(b.scale2)
    return b


def pack_block2d(sd, prefix, cin, cout, split):
    b = Block()
    cp = pad64(cout)
    b.cin, b.cmid, b.taps, b.pad, b.lc = cin, cp, (1, 3, 3), (0, 1, 1), cout
    w_main = sd[prefix + ".main.0.weight"].unsqueeze(2)
    w_ds = sd[prefix + ".downsample.0.weight"].unsqueeze(2)
    b.w1 = pack_conv([w_main, w_ds], cin, [cp, cp], split)
    dev = w_main.device
    b.scale1 = b.shift1 = b.scale2 = b.shift2 = None
    b.slope1 = torch.cat([sd[prefix + ".main.1.weight"].float().expand(cp), torch.ones(cp, device=dev)]).contiguous()   # PReLU | identity
    b.w2 = pack_conv([sd[prefix + ".main.2.weight"].unsqueeze(2)], cp, [cp], split)
    b.slope2 = sd[prefix + ".relu.weight"].float().expand(cp).contiguous()
    return b


def block_wants_q(x, blk, tmp):
    """True when the block's first convolution would read ``x`` through the two-unit kernel (ops.quant()): the producer of ``x`` then
    writes the quantised operand planes from its epilogue (``out_q``) instead of leaving them to a separate pass."""
    return ops._QUANT and ops.conv_gemm(x, blk.cin, blk.w1, 2 * blk.cmid, kernel=blk.taps, pad=blk.pad, out=tmp, probe=True)


def run_block(x, blk, tmp, out, o_ch_off=0, out_q=False):
    """x -> out[..., o_ch_off : o_ch_off + cmid];  tmp is a ``[..., 2*cmid]`` scratch tensor."""
    c = blk.cmid
    q2 = ops._QUANT and ops.conv_gemm(tmp, c, blk.w2, c, kernel=blk.taps, pad=blk.pad, out=out, o_ch_off=o_ch_off, probe=True)
    ops.conv_gemm(x, blk.cin, blk.w1, 2 * c, kernel=blk.taps, pad=blk.pad, scale=blk.scale1, shift=blk.shift1,
                  slope=blk.slope1, out=tmp, lcout=2 * blk.lc, out_q=q2)
    ops.conv_gemm(tmp, c, blk.w2, c, kernel=blk.taps, pad=blk.pad, scale=blk.scale2, shift=blk.shift2, slope=blk.slope2,
                  residual=tmp, r_ch_off=c, out=out, o_ch_off=o_ch_off, lcin=blk.lc, lcout=blk.lc, out_q=out_q)
    return out


# ---------------------------------------------------------------------------------------------- encoder (layers.py:186-217)
class EncoderWeights(object):
    def __init__(self, sd, prefix, nf, group_frames, split):
        self.l1_w = pack_conv([sd[prefix + ".layer1.0.weight"]], pad64(nf), [2 * nf], split)
        self.l1_shift = sd[prefix + ".layer1.0.bias"].float().contiguous()
        self.blocks = [pack_block3d(sd, prefix + ".layer1.1", 2 * nf, 2 * nf, split),
                       pack_block3d(sd, prefix + ".layer2.1", 2 * nf, 4 * nf, split),
                       pack_block3d(sd, prefix + ".layer2.2", 4 * nf, 4 * nf, split),
                       pack_block3d(sd, prefix + ".layer3.1", 4 * nf, 8 * nf, split),
                       pack_block3d(sd, prefix + ".layer3.2", 8 * nf, 8 * nf, split)]
        self.merge = [pack_conv([sd[prefix + ".l1temporalMerge.weight"]], 2 * nf, [2 * nf], split),
                      pack_conv([sd[prefix + ".l2temporalMerge.weight"]], 4 * nf, [4 * nf], split),
                      pack_conv([sd[prefix + ".temporalMerge.weight"]], 8 * nf, [8 * nf], split)]
        self.nf, self.g = nf, group_frames


class EncoderBuffers(object):
    def __init__(self, b, nf, g, dev, split):
        mk = lambda *shape: SplitTensor.empty(shape, dev, split)
        self.l1a = mk(b, g, 64, 64, 2 * nf)
        self.t1 = mk(b, g, 64, 64, 4 * nf)
        self.l1 = mk(b, g, 64, 64, 2 * nf)
        self.l2in = mk(b, g // 2, 32, 32, 2 * nf)
        self.t2 = mk(b, g // 2, 32, 32, 8 * nf)
        self.l2a = mk(b, g // 2, 32, 32, 4 * nf)
        self.l2 = mk(b, g // 2, 32, 32, 4 * nf)
        self.l3in = mk(b, g // 4, 16, 16, 4 * nf)
        self.t3 = mk(b, g // 4, 16, 16, 16 * nf)
        self.l3a = mk(b, g // 4, 16, 16, 8 * nf)
        self.l3 = mk(b, g // 4, 16, 16, 8 * nf)
        self.f1 = mk(b, 1, 64, 64, 2 * nf)
        self.f2 = mk(b, 1, 32, 32, 4 * nf)
        self.f3 = mk(b, 1, 16, 16, 8 * nf)


def run_encoder(x, w, bf):
    """x: SplitTensor ``[B, G, 64, 64, nf]`` chirp features -> (f1, f2, f3) temporal-merged maps."""
    nf, g = w.nf, w.g
    ops.conv_gemm(x, pad64(nf), w.l1_w, 2 * nf, kernel=(3, 3, 3), pad=(1, 1, 1), shift=w.l1_shift, out=bf.l1a,
                  out_q=block_wants_q(bf.l1a, w.blocks[0], bf.t1))
    run_block(bf.l1a, w.blocks[0], bf.t1, bf.l1)
    ops.resample_linear(bf.l1, 2 * nf, bf.l2in)
    run_block(bf.l2in, w.blocks[1], bf.t2, bf.l2a, out_q=block_wants_q(bf.l2a, w.blocks[2], bf.t2))
    run_block(bf.l2a, w.blocks[2], bf.t2, bf.l2)
    ops.resample_linear(bf.l2, 4 * nf, bf.l3in)
    run_block(bf.l3in, w.blocks[3], bf.t3, bf.l3a, out_q=block_wants_q(bf.l3a, w.blocks[4], bf.t3))
    run_block(bf.l3a, w.blocks[4], bf.t3, bf.l3)
    ops.conv_gemm(bf.l1, 2 * nf, w.merge[0], 2 * nf, kernel=(g, 1, 1), out=bf.f1)
    ops.conv_gemm(bf.l2, 4 * nf, w.merge[1], 4 * nf, kernel=(g // 2, 1, 1), out=bf.f2)
    ops.conv_gemm(bf.l3, 8 * nf, w.merge[2], 8 * nf, kernel=(g // 4, 1, 1), out=bf.f3)
    return bf.f1, bf.f2, bf.f3


# ---------------------------------------------------------------------------------------------- decoder (layers.py:72-184)
PROJ_HORI = ("phi_cross_hori", "theta_cross_hori", "phi_self_hori", "theta_self_hori")
PROJ_VERT = ("phi_cross_vert", "theta_cross_vert", "phi_self_vert", "theta_self_vert")


class DecoderWeights(object):
    def __init__(self, sd, nf, keypoints, split):
        p = "radarDecoder."
        self.nf, self.keypoints = nf, keypoints
        self.proj_hori, self.proj_vert = [], []
        for level, c in enumerate((8 * nf, 4 * nf, 2 * nf)):
            for names, dst in ((PROJ_HORI, self.proj_hori), (PROJ_VERT, self.proj_vert)):
                ws = [sd[p + "%s.%d.weight" % (n, level)].unsqueeze(2) for n in names]
                dst.append(pack_conv(ws, c, [c] * 4, split))
        self.blocks = [pack_block2d(sd, p + "decoderLayer3.0", 32 * nf, 8 * nf, split),
                       pack_block2d(sd, p + "decoderLayer3.1", 8 * nf, 4 * nf, split),
                       pack_block2d(sd, p + "decoderLayer2.0", 20 * nf, 4 * nf, split),
                       pack_block2d(sd, p + "decoderLayer2.1", 4 * nf, 2 * nf, split),
                       pack_block2d(sd, p + "decoderLayer1.0", 10 * nf, 2 * nf, split),
                       pack_block2d(sd, p + "decoderLayer1.1", 2 * nf, nf, split)]
        self.head = pack_conv([sd[p + "decoderLayer1.2.weight"].unsqueeze(2)], pad64(nf), [pad64(keypoints)], split)
        self.gcn_w = [sd[p + "gcn.L%d.weight" % i].float().contiguous() for i in (1, 2, 3)]
        self.gcn_b = [sd[p + "gcn.L%d.bias" % i].float().contiguous() for i in (1, 2, 3)]
        # tensor-core PRGCN: W[q, p] is already the K-major B operand of  Yt[(b,j), q] = sum_p St[(b,j), p] W[q, p]
        self.gcn_w_tc = [SplitTensor.from_float(w.view(1, 1024, 1024), True) for w in self.gcn_w] if split else None
        self.gcn_relu = torch.zeros(1024, dtype=torch.float32, device=self.gcn_w[0].device)


class DecoderBuffers(object):
    def __init__(self, b, nf, keypoints, dev, split):
        mk = lambda *shape: SplitTensor.empty(shape, dev, split)
        self.levels = []
        for c, hw, prev in ((8 * nf, 16, 0), (4 * nf, 32, 4 * nf), (2 * nf, 64, 2 * nf)):
            s = hw * hw
            self.levels.append(dict(c=c, hw=hw, s=s, prev=prev,
                                    proj_ra=mk(b, 1, 1, s, 4 * c), proj_re=mk(b, 1, 1, s, 4 * c),
                                    vt_ra=mk(b, c, s), vt_re=mk(b, c, s),
                                    cat=mk(b, 1, hw, hw, prev + 4 * c)))
        smax = 64 * 64 if not split else 16 * 16      # levels 1 and 2 run fused when split; scratch covers the unfused level(s)
        self._scratch_elems, self._split, self._dev = b * smax * smax, split, dev
        self._scratch = {}                             # index -> (fp32 logits, split probabilities) of one unfused attention
        self._att_streams = None
        self.attention_scratch(0)
        self.t3a, self.o3a = mk(b, 1, 16, 16, 16 * nf), mk(b, 1, 16, 16, 8 * nf)
        self.t3b, self.o3b = mk(b, 1, 16, 16, 8 * nf), mk(b, 1, 16, 16, 4 * nf)
        self.t2a, self.o2a = mk(b, 1, 32, 32, 8 * nf), mk(b, 1, 32, 32, 4 * nf)
        self.t2b, self.o2b = mk(b, 1, 32, 32, 2 * pad64(2 * nf)), mk(b, 1, 32, 32, pad64(2 * nf))
        self.t1a, self.o1a = mk(b, 1, 64, 64, 2 * pad64(2 * nf)), mk(b, 1, 64, 64, pad64(2 * nf))
        self.t1b, self.o1b = mk(b, 1, 64, 64, 2 * pad64(nf)), mk(b, 1, 64, 64, pad64(nf))
        self.logits_out = torch.empty((b, 64 * 64, pad64(keypoints)), dtype=torch.float32, device=dev)
        self.gcn_ws = torch.empty(max(ops.prgcn_workspace_bytes(b) // 4, 4), dtype=torch.float32, device=dev)
        self.gcn_rows = -(-(b * keypoints) // 128) * 128          # rows (b, j) of the transposed GCN activations, padded to the GEMM tile
        if split:
            self.gcn_st = [SplitTensor.empty((1, 1, 1, self.gcn_rows, 1024), dev, True, zero=True) for _ in range(2)]
            self.gcn_y3 = torch.zeros((1, 1, 1, self.gcn_rows, 1024), dtype=torch.float32, device=dev)
            self.gcn_bias_rows = None
        self.heatmap = torch.empty((b, keypoints, 64, 64), dtype=torch.float32, device=dev)
        self.gcn_heatmap = torch.empty((b, keypoints, 64, 64), dtype=torch.float32, device=dev)

    def attention_scratch(self, index):
        """(fp32 logits, split probabilities) scratch of unfused attention number ``index`` of a level."""
        pair = self._scratch.get(index)
        if pair is None:
            pair = (torch.empty(self._scratch_elems, dtype=torch.float32, device=self._dev),
                    SplitTensor.empty((self._scratch_elems,), self._dev, self._split))
            self._scratch[index] = pair
        return pair

    def attention_streams(self, device):
        """Three side streams on which attentions 1..3 of a level run next to attention 0 (small batches)."""
        if self._att_streams is None:
            self._att_streams = [torch.cuda.Stream(device=device) for _ in range(3)]
        return self._att_streams


def _view(t, shape):
    n = 1
    for d in shape:
        n *= d
    return SplitTensor(t.hi[:n].view(shape), None if t.lo is None else t.lo[:n].view(shape))


def run_attention(bf, lv, q_src, q_off, k_src, k_off, v, vt, out, o_off, residual, scratch=0):
    """layers.py:126-133 with Q = q_src[..., q_off:+C], K = k_src[..., k_off:+C], V = v:
    logits[n, m] = <Q[n], K[m]>;  P = softmax over keys m;  out[n] = sum_m P[n, m] V[m]  (+ V[n] for the cross branches).
    ``scratch``: which logits / probabilities scratch pair the unfused path uses (attentions running concurrently need their own)."""
    b = v.hi.shape[0]
    c, s = lv["c"], lv["s"]
    if c in (64, 128) and v.lo is not None:      # fused flash-style kernel (levels 1 and 2: 99.7 % of the attention FLOPs)
        ops.attention_fwd(q_src, q_off, k_src, k_off, vt, c, out, o_off, residual=v if residual else None)
        return
    lg, pb = bf.attention_scratch(scratch)
    logits = lg[:b * s * s].view(b, 1, 1, s, s)
    probs = _view(pb, (b, 1, 1, s, s))
    ops.conv_gemm(q_src, c, SplitTensor(k_src.hi.view(b, s, 4 * c), None if k_src.lo is None else k_src.lo.view(b, s, 4 * c)), s,
                  a_ch_off=q_off, w_batched=True, w_ld=4 * c, w_ch_off=k_off, out_f32=logits)
    ops.softmax_rows(logits, probs)
    ops.conv_gemm(probs, s, vt, c, w_batched=True, residual=v if residual else None, out=out, o_ch_off=o_off)


def run_attention_level(bf, lv, w_hori, w_vert, ra, re, concurrent=False):
    """One scale of the cross/self attention (layers.py:138-149): fills cat[..., prev : prev + 4C] with
    (ra_cross, ra_self, re_cross, re_self).  ``concurrent`` (small batches): the four attentions are independent and each fills only
    batch * S / 128 CTAs (32 of 148 SMs at batch 1, level 1), so they run as four parallel branches (fork / join on side streams, parallel
    branches of the captured graph) instead of back to back."""
    c, prev = lv["c"], lv["prev"]
    b, hw = ra.hi.shape[0], lv["hw"]
    ra_s = SplitTensor(ra.hi.view(b, 1, 1, lv["s"], c), None if ra.lo is None else ra.lo.view(b, 1, 1, lv["s"], c))
    re_s = SplitTensor(re.hi.view(b, 1, 1, lv["s"], c), None if re.lo is None else re.lo.view(b, 1, 1, lv["s"], c))
    ops.conv_gemm(ra_s, c, w_hori, 4 * c, out=lv["proj_ra"])     # channels: phi_c_hori | theta_c_hori | phi_s_hori | theta_s_hori
    ops.conv_gemm(re_s, c, w_vert, 4 * c, out=lv["proj_re"])     #           phi_c_vert | theta_c_vert | phi_s_vert | theta_s_vert
    ops.transpose_split(ra_s, c, lv["vt_ra"])
    ops.transpose_split(re_s, c, lv["vt_re"])
    cat = lv["cat"]
    cat_s = SplitTensor(cat.hi.view(b, 1, 1, lv["s"], prev + 4 * c), None if cat.lo is None else cat.lo.view(b, 1, 1, lv["s"], prev + 4 * c))
    pra, pre = lv["proj_ra"], lv["proj_re"]
    jobs = [
        (pre, c, pra, 0, ra_s, lv["vt_ra"], prev, True),                 # ra_cross: k = phi_c_hori(ra), q = theta_c_vert(re), v = ra, + ra
        (pra, 3 * c, pra, 2 * c, ra_s, lv["vt_ra"], prev + c, False),    # ra_self : k = phi_s_hori(ra), q = theta_s_hori(ra), v = ra
        (pra, c, pre, 0, re_s, lv["vt_re"], prev + 2 * c, True),         # re_cross: k = phi_c_vert(re), q = theta_c_hori(ra), v = re, + re
        (pre, 3 * c, pre, 2 * c, re_s, lv["vt_re"], prev + 3 * c, False),   # re_self : k = phi_s_vert(re), q = theta_s_vert(re), v = re
    ]
    if not concurrent:
        for q_src, q_off, k_src, k_off, v, vt, o_off, res in jobs:
            run_attention(bf, lv, q_src, q_off, k_src, k_off, v, vt, cat_s, o_off, res)
        return cat
    main = torch.cuda.current_stream()
    sides = bf.attention_streams(ra.hi.device)
    for side in sides:
        side.wait_stream(main)                                           # fork: projections and transposes are done on the main stream
    for i, (q_src, q_off, k_src, k_off, v, vt, o_off, res) in enumerate(jobs):
        if i == 0:
            run_attention(bf, lv, q_src, q_off, k_src, k_off, v, vt, cat_s, o_off, res, scratch=0)
        else:
            with torch.cuda.stream(sides[i - 1]), ops.coop_slot(("attention", i)):
                run_attention(bf, lv, q_src, q_off, k_src, k_off, v, vt, cat_s, o_off, res, scratch=i)
    for side in sides:
        main.wait_stream(side)                                           # join
    return cat


def run_decoder(w, bf, feats_ra, feats_re, adj, concurrent=False):
    """feats_* = (f1, f2, f3).  Returns (heatmap, gcn_heatmap) float32 ``[B, 14, 64, 64]``."""
    nf = w.nf
    l3, l2, l1 = bf.levels
    cat3 = run_attention_level(bf, l3, w.proj_hori[0], w.proj_vert[0], feats_ra[2], feats_re[2], concurrent)
    run_block(cat3, w.blocks[0], bf.t3a, bf.o3a, out_q=block_wants_q(bf.o3a, w.blocks[1], bf.t3b))
    run_block(bf.o3a, w.blocks[1], bf.t3b, bf.o3b)
    ops.resample_linear(bf.o3b, 4 * nf, l2["cat"])
    cat2 = run_attention_level(bf, l2, w.proj_hori[1], w.proj_vert[1], feats_ra[1], feats_re[1], concurrent)
    run_block(cat2, w.blocks[2], bf.t2a, bf.o2a, out_q=block_wants_q(bf.o2a, w.blocks[3], bf.t2b))
    run_block(bf.o2a, w.blocks[3], bf.t2b, bf.o2b)
    ops.resample_linear(bf.o2b, 2 * nf, l1["cat"])
    cat1 = run_attention_level(bf, l1, w.proj_hori[2], w.proj_vert[2], feats_ra[0], feats_re[0], concurrent)
    run_block(cat1, w.blocks[4], bf.t1a, bf.o1a, out_q=block_wants_q(bf.o1a, w.blocks[5], bf.t1b))
    run_block(bf.o1a, w.blocks[5], bf.t1b, bf.o1b)
    b = bf.o1b.hi.shape[0]
    ops.conv_gemm(bf.o1b, pad64(nf), w.head, bf.logits_out.shape[-1], out_f32=bf.logits_out.view(b, 1, 64, 64, -1), lcin=nf, lcout=w.keypoints)
    if w.gcn_w_tc is None:
        ops.prgcn_fwd(bf.logits_out, w.gcn_w, w.gcn_b, adj, bf.gcn_ws, bf.heatmap, bf.gcn_heatmap)
        return bf.heatmap, bf.gcn_heatmap
    if bf.gcn_bias_rows is None:      # bias[q, j] as residual rows [(b, j)][q], built once per plan
        rows = []
        for bias in w.gcn_b:
            t = torch.zeros(bf.gcn_rows, 1024, dtype=torch.float32, device=bias.device)
            t[:b * bias.shape[1]] = bias.t().repeat(b, 1)
            rows.append(SplitTensor.from_float(t.view(1, 1, 1, bf.gcn_rows, 1024), True))
        bf.gcn_bias_rows = rows
    st_a, st_b = bf.gcn_st
    ops.gcn_nodes(bf.logits_out, adj, bf.heatmap, st_a)
    lr = b * w.keypoints                 # GEMM rows that carry data (the rest is tile padding)
    ops.conv_gemm(st_a, 1024, w.gcn_w_tc[0], 1024, residual=bf.gcn_bias_rows[0], slope=w.gcn_relu, out=st_b, lrows=lr)
    ops.gcn_mix(st_b, adj, st_a, b)
    ops.conv_gemm(st_a, 1024, w.gcn_w_tc[1], 1024, residual=bf.gcn_bias_rows[1], slope=w.gcn_relu, out=st_b, lrows=lr)
    ops.gcn_mix(st_b, adj, st_a, b)
    ops.conv_gemm(st_a, 1024, w.gcn_w_tc[2], 1024, residual=bf.gcn_bias_rows[2], out_f32=bf.gcn_y3, lrows=lr)
    ops.gcn_heads(bf.gcn_y3.view(bf.gcn_rows, 1024), bf.gcn_heatmap, b)
    return bf.heatmap, bf.gcn_heatmap
