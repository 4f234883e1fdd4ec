"""HuPRNet on the B200: host-side mirror of /root/reference/models/networks.py (class ``HuPRNet``).

Same constructor (``HuPRNet(cfg)``), same ``forward(VRDAEmaps_hori, VRDAEmaps_vert) -> (heatmap [B,14,1,64,64],
gcn_heatmap [B,1,14,64,64])`` and the same 255-entry ``state_dict`` (names, shapes, construction-order initialisation,
networks.py:17-21), so reference checkpoints load unchanged.  The arithmetic is NOT torch: every stage is a launch into
libhupr_b200.so (MNet kernel, tcgen05 implicit-GEMM convolutions / attention matmuls, resampling, softmax, fused PRGCN).
There is no CPU path — calling ``forward`` without a B200 raises.

``model.eval()`` (BatchNorm uses running statistics, reference tools/run.py:36) runs the planned inference launch sequence below.
``model.train()`` routes ``forward`` through ``hupr_b200.training.TrainStep`` (batch statistics, running-stat updates) and returns
outputs that are connected to the parameters by a ``torch.autograd.Function`` whose backward is TrainStep's hand-written backward
pass — so the reference loop ``preds = model(h, v); loss.backward(); optimizer.step()`` (tools/run.py:76-79) drives this module with
any torch optimiser.  The fused path (``TrainStep.forward_backward`` + ``hupr_adam_step``, one CUDA graph) is what ``Runner`` and
``bench.py`` use.
"""
import math

import os

import torch
import torch.nn as nn

from .. import ops
from ..ops import SplitTensor
from . import layers as L

ADJACENCY = (   # models/layers.py:97-112 — asymmetric and un-normalised on purpose (trained weights depend on it)
    (1, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0), (1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0), (0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0),
    (1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0), (0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0), (0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0),
    (0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0), (0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0), (0, 0, 0, 0, 0, 0, 1, 0, 1, 1, 0, 0, 0, 0),
    (0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0), (0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0), (0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0),
    (0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1), (0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1))


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's dotted parameter names."""


def _register(root, name, tensor, kind):
    parts = name.split(".")
    node = root
    for part in parts[:-1]:
        if part not in node._modules:
            node.add_module(part, _Node())
        node = node._modules[part]
    if kind == "param":
        node.register_parameter(parts[-1], nn.Parameter(tensor))
    else:
        node.register_buffer(parts[-1], tensor)


def _conv(root, name, cout, cin, kernel, bias):
    w = torch.empty((cout, cin) + tuple(kernel))
    nn.init.kaiming_uniform_(w, a=math.sqrt(5))          # torch's Conv default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    _register(root, name + ".weight", w, "param")
    if bias:
        fan_in = cin
        for k in kernel:
            fan_in *= k
        b = torch.empty(cout)
        nn.init.uniform_(b, -1.0 / math.sqrt(fan_in), 1.0 / math.sqrt(fan_in))
        _register(root, name + ".bias", b, "param")


def _bn(root, name, c):
    _register(root, name + ".weight", torch.ones(c), "param")
    _register(root, name + ".bias", torch.zeros(c), "param")
    _register(root, name + ".running_mean", torch.zeros(c), "buffer")
    _register(root, name + ".running_var", torch.ones(c), "buffer")
    _register(root, name + ".num_batches_tracked", torch.tensor(0, dtype=torch.long), "buffer")


def _block3d(root, name, cin, cout):
    _conv(root, name + ".main.0", cout, cin, (3, 3, 3), False)
    _bn(root, name + ".main.1", cout)
    _conv(root, name + ".main.3", cout, cout, (3, 3, 3), False)
    _bn(root, name + ".main.4", cout)
    _conv(root, name + ".downsample.0", cout, cin, (3, 3, 3), False)
    _bn(root, name + ".downsample.1", cout)


def _block2d(root, name, cin, cout):
    _conv(root, name + ".main.0", cout, cin, (3, 3), False)
    _register(root, name + ".main.1.weight", torch.full((1,), 0.25), "param")       # nn.PReLU default
    _conv(root, name + ".main.2", cout, cout, (3, 3), False)
    _conv(root, name + ".downsample.0", cout, cin, (3, 3), False)
    _register(root, name + ".relu.weight", torch.full((1,), 0.25), "param")


class _TrainForward(torch.autograd.Function):
    """Autograd bridge of the train-mode forward: forward = TrainStep.forward, backward = TrainStep.backward (every parameter gradient
    computed by the library in one pass), handed to autograd as the gradients of the parameter inputs."""

    @staticmethod
    def forward(ctx, step, hori, vert, *params):
        heat, gcn = step.forward(hori, vert)
        ctx.step = step
        ctx.names = [n for n, _ in step.model.named_parameters()]
        return heat, gcn

    @staticmethod
    def backward(ctx, g_heat, g_gcn):
        step = ctx.step
        step.backward(g_heat, g_gcn)
        grads = tuple(step.gview[n].clone() for n in ctx.names)      # the flat gradient buffer is overwritten by the next backward
        return (None, None, None) + grads


class HuPRNet(nn.Module):
    def __init__(self, cfg, split=True):
        """``split=True``: bf16 hi/lo activations and weights, three tensor-core products per k-step — fp32-equivalent
        results (the reference computes in fp32).  ``split=False``: single bf16 product (faster, ~1e-2 accuracy).

        ``quant_cross_terms`` (attribute, default True unless HUPR_QUANT=0; eval-mode ``split`` forward only): the large 3-tap
        convolutions (cout a multiple of 128, enough tiles to fill the GPU) evaluate the two hi*lo cross terms as e4m3 products next to an
        fp16 main product (ops.quant: two tensor units per k-step instead of three).  Heat maps stay well inside the 1e-3
        bar (tests/test_model_gpu.py, tests/test_pipeline_gpu.py: the error against the oracle is the same size as with three bf16
        products), keypoints unchanged.  Range: activations keep full accuracy for |x| < 224 and must stay below 16 376 (fp16 plane);
        filters with a weight >= 3.99 keep the three bf16 products automatically.  Set False for three bf16 products everywhere."""
        super(HuPRNet, self).__init__()
        self.quant_cross_terms = os.environ.get("HUPR_QUANT", "1") != "0"
        self.numFrames = cfg.DATASET.numFrames
        self.numFilters = nf = cfg.MODEL.numFilters
        self.rangeSize = cfg.DATASET.rangeSize
        self.heatmapSize = cfg.DATASET.heatmapSize
        self.azimuthSize = cfg.DATASET.azimuthSize
        self.elevationSize = cfg.DATASET.elevationSize
        self.numGroupFrames = g = cfg.DATASET.numGroupFrames
        self.numKeypoints = kp = cfg.DATASET.numKeypoints
        if (self.numFrames, nf, self.rangeSize, self.azimuthSize, self.elevationSize, g, kp, self.heatmapSize) != (8, 32, 64, 64, 8, 8, 14, 64):
            raise ValueError("hupr_b200 kernels are specialised for config/mscsa_prgcn.yaml "
                             "(numFrames 8, numFilters 32, 64x64x8 cubes, 8 group frames, 14 keypoints)")
        self.split = split
        # parameters, in the reference's construction order (networks.py:17-21) so that seeded initialisation matches
        for net in ("RAchirpNet", "REchirpNet"):
            _conv(self, net + ".temporalConvWx1x1", nf, 2, (2, 1, 1), True)
        for enc in ("RAradarEncoder", "REradarEncoder"):
            _conv(self, enc + ".layer1.0", 2 * nf, nf, (3, 3, 3), True)
            _block3d(self, enc + ".layer1.1", 2 * nf, 2 * nf)
            _block3d(self, enc + ".layer2.1", 2 * nf, 4 * nf)
            _block3d(self, enc + ".layer2.2", 4 * nf, 4 * nf)
            _block3d(self, enc + ".layer3.1", 4 * nf, 8 * nf)
            _block3d(self, enc + ".layer3.2", 8 * nf, 8 * nf)
            _conv(self, enc + ".l1temporalMerge", 2 * nf, 2 * nf, (g, 1, 1), False)
            _conv(self, enc + ".l2temporalMerge", 4 * nf, 4 * nf, (g // 2, 1, 1), False)
            _conv(self, enc + ".temporalMerge", 8 * nf, 8 * nf, (g // 4, 1, 1), False)
        dec = "radarDecoder"
        _block2d(self, dec + ".decoderLayer3.0", 32 * nf, 8 * nf)
        _block2d(self, dec + ".decoderLayer3.1", 8 * nf, 4 * nf)
        _block2d(self, dec + ".decoderLayer2.0", 20 * nf, 4 * nf)
        _block2d(self, dec + ".decoderLayer2.1", 4 * nf, 2 * nf)
        _block2d(self, dec + ".decoderLayer1.0", 10 * nf, 2 * nf)
        _block2d(self, dec + ".decoderLayer1.1", 2 * nf, nf)
        _conv(self, dec + ".decoderLayer1.2", kp, nf, (1, 1), False)
        feat = (self.heatmapSize // 2) ** 2
        for layer in ("L1", "L2", "L3"):
            stdv = 1.0 / math.sqrt(feat)                 # gcn_networks.py:17-21
            _register(self, dec + ".gcn.%s.weight" % layer, torch.empty(feat, feat).uniform_(-stdv, stdv), "param")
            _register(self, dec + ".gcn.%s.bias" % layer, torch.empty(feat, kp).uniform_(-stdv, stdv), "param")
        for name in ("phi_cross_hori", "theta_cross_hori", "phi_cross_vert", "theta_cross_vert",
                     "phi_self_hori", "theta_self_hori", "phi_self_vert", "theta_self_vert"):
            for i, c in enumerate((8 * nf, 4 * nf, 2 * nf)):
                _conv(self, dec + ".%s.%d" % (name, i), c, c, (1, 1), False)
        self._packed = None
        self._plans = {}
        self._train_step = None      # the TrainStep that owns this model's flat parameter buffer (set by TrainStep.__init__)

    # ------------------------------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict, strict=True, **kwargs):
        out = super(HuPRNet, self).load_state_dict(state_dict, strict, **kwargs)
        self.invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super(HuPRNet, self)._apply(fn, *args, **kwargs)
        self.invalidate()
        self._train_step = None      # parameter storage was replaced: a TrainStep's flat views are stale
        return out

    def __getstate__(self):
        """Pickling / deepcopy keep parameters and buffers only: packed operands, launch plans and the training step hold device scratch,
        streams and raw-pointer tables that belong to this process."""
        state = self.__dict__.copy()
        state["_packed"], state["_plans"], state["_train_step"] = None, {}, None
        return state

    def quant_saturations(self, reset=True):
        """Values that left the range of the two-unit convolutions' operand planes (|activation| >= 16 376) on this model's device since
        the last reset — 0 in normal operation (ops.quant_saturations; synchronises).  A non-zero count means those forwards were
        wrong: set ``quant_cross_terms = False`` for this checkpoint / input scaling."""
        return ops.quant_saturations(next(self.parameters()).device, reset)

    def invalidate(self):
        """Drop the packed (kernel-format) weights and cached launch plans; call after mutating parameters in place."""
        self._packed = None
        self._plans = {}

    def _pack(self):
        sd = self.state_dict()
        dev = sd["RAchirpNet.temporalConvWx1x1.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("hupr_b200.HuPRNet runs on a B200 only: move the module to a CUDA device (no CPU fallback)")
        nf, g, kp = self.numFilters, self.numGroupFrames, self.numKeypoints
        pk = {"device": dev}
        for net in ("RAchirpNet", "REchirpNet"):
            pk[net] = (sd[net + ".temporalConvWx1x1.weight"].float().contiguous(), sd[net + ".temporalConvWx1x1.bias"].float().contiguous())
        for enc in ("RAradarEncoder", "REradarEncoder"):
            pk[enc] = L.EncoderWeights(sd, enc, nf, g, self.split)
        pk["decoder"] = L.DecoderWeights(sd, nf, kp, self.split)
        pk["adj"] = torch.tensor(ADJACENCY, dtype=torch.float32, device=dev)
        self._packed = pk
        return pk

    def _plan(self, batch):
        pk = self._packed or self._pack()
        plan = self._plans.get(batch)
        if plan is None:
            dev, nf, g = pk["device"], self.numFilters, self.numGroupFrames
            plan = {
                "chirp_ra": SplitTensor.empty((batch, g, 64, 64, nf), dev, self.split),
                "chirp_re": SplitTensor.empty((batch, g, 64, 64, nf), dev, self.split),
                "enc_ra": L.EncoderBuffers(batch, nf, g, dev, self.split),
                "enc_re": L.EncoderBuffers(batch, nf, g, dev, self.split),
                "dec": L.DecoderBuffers(batch, nf, self.numKeypoints, dev, self.split),
                "coop": ops.CoopWorkspaces(dev),      # split-K scratch of this plan's launch sequence (slot 1 = side stream)
            }
            self._plans = {batch: plan}      # one live plan: the buffers are sized by batch
        return pk, plan

    def chirp_weights(self):
        """(RAchirpNet weight, bias, REchirpNet weight, bias) as contiguous float32 device tensors (the MNet kernels' inputs)."""
        pk = self._packed or self._pack()
        return pk["RAchirpNet"] + pk["REchirpNet"]

    # ------------------------------------------------------------------------------------------ forward
    def forward_chirp(self, VRDAEmaps_hori, VRDAEmaps_vert):
        """networks.py:23-33.  Returns channels-last split tensors ``[B, G, 64, 64, numFilters]`` (the reference returns the
        same values as ``[B, numFilters, G, 64, 64]`` fp32)."""
        batch = VRDAEmaps_hori.size(0)
        pk, plan = self._plan(batch)
        for x, net, key in ((VRDAEmaps_hori, "RAchirpNet", "chirp_ra"), (VRDAEmaps_vert, "REchirpNet", "chirp_re")):
            if not x.is_cuda or x.dtype != torch.float32:
                raise TypeError("HuPRNet.forward expects float32 CUDA tensors [B, 8, 8, 2, 64, 64, 8]")
            if tuple(x.shape[1:]) != (self.numGroupFrames, self.numFrames, 2, self.rangeSize, self.azimuthSize, self.elevationSize):
                raise ValueError("unexpected VRDAE shape %s" % (tuple(x.shape),))
            w, b = pk[net]
            ops.mnet_fwd(x.contiguous(), w, b, plan[key])
        return plan["chirp_ra"], plan["chirp_re"]

    def forward_features(self, chirp_ra, chirp_re):
        """Encoders + decoder + PRGCN on chirp features.  Returns float32 ``[B,14,64,64]`` heatmap and gcn heatmap (views of the
        plan's output buffers — valid until the next forward)."""
        batch = chirp_ra.hi.shape[0]
        pk, plan = self._plan(batch)
        with ops.coop_scope(plan["coop"]), ops.pdl(batch <= 4 or ops._C.lib().hupr_set_pdl(-1) == 1), \
                ops.quant(self.split and self.quant_cross_terms):
            return self._forward_features(pk, plan, batch, chirp_ra, chirp_re)

    def _forward_features(self, pk, plan, batch, chirp_ra, chirp_re):
        small = batch <= 4 and ops._PROFILE is None
        if small:
            # small batches leave most SMs idle inside one encoder (level 2/3 convolutions launch 16-64 CTAs): run the two
            # independent sensor branches on two streams (fork/join with events, captured into the CUDA graph as parallel branches)
            main = torch.cuda.current_stream()
            side = plan.get("side_stream")
            if side is None:
                side = plan["side_stream"] = torch.cuda.Stream(device=pk["device"])
            side.wait_stream(main)
            with torch.cuda.stream(side), ops.coop_slot(1):
                feats_re = L.run_encoder(chirp_re, pk["REradarEncoder"], plan["enc_re"])
            feats_ra = L.run_encoder(chirp_ra, pk["RAradarEncoder"], plan["enc_ra"])
            main.wait_stream(side)
        else:
            feats_ra = L.run_encoder(chirp_ra, pk["RAradarEncoder"], plan["enc_ra"])
            feats_re = L.run_encoder(chirp_re, pk["REradarEncoder"], plan["enc_re"])
        return L.run_decoder(pk["decoder"], plan["dec"], feats_ra, feats_re, pk["adj"], concurrent=small)

    def _forward_train(self, hori, vert):
        """networks.py:35-41 under ``model.train()`` (tools/run.py:66,76): batch-statistics BatchNorm, outputs connected to autograd."""
        from ..training import TrainStep
        if next(self.parameters()).device.type != "cuda" or not hori.is_cuda:
            raise RuntimeError("hupr_b200.HuPRNet runs on a B200 only: move the module and its inputs to a CUDA device (no CPU fallback)")
        if not self.split:
            raise RuntimeError("train-mode forward needs the hi/lo (split=True) model")
        step = self._train_step
        if step is None:
            step = TrainStep(self, install_grads=False)       # registers itself as self._train_step
        for q in self.parameters():
            if q.grad is not None and q.grad.data_ptr() == step.gview_ptr(q):
                q.grad = None            # gradients installed by the fused path alias the flat buffer: autograd must own .grad here
        step._dirty = True               # any optimiser may have updated the parameters in place since the last pass
        heat, gcn = _TrainForward.apply(step, hori.float().contiguous(), vert.float().contiguous(), *self.parameters())
        return heat.unsqueeze(2), gcn.unsqueeze(1)

    def forward(self, VRDAEmaps_hori, VRDAEmaps_vert):
        if self.training:
            return self._forward_train(VRDAEmaps_hori, VRDAEmaps_vert)
        ra, re = self.forward_chirp(VRDAEmaps_hori, VRDAEmaps_vert)
        heatmap, gcn = self.forward_features(ra, re)
        return heatmap.unsqueeze(2), gcn.unsqueeze(1)
