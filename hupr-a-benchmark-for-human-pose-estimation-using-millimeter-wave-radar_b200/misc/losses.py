"""Heatmap loss on the B200: host-side mirror of /root/reference/misc/losses.py (``LossComputer``), misc/utils.py
(``generateTarget``) and misc/metrics.py (``get_max_preds``).

``LossComputer(cfg, device).computeLoss(preds, gt) -> (loss, loss2, pred2d, gt2d)`` keeps the reference's signature and return
convention (losses.py:23-45): ``loss``/``loss2`` are 0-dim float32 tensors on the device, ``pred2d``/``gt2d`` are float32 numpy
arrays ``[B, 14, 2]`` of heatmap-pixel (x, y).  The reference synthesises the Gaussian targets on the CPU, copies them to the GPU
and copies the heatmaps back for a numpy argmax; here one kernel pair (``hupr_heatmap_loss_fwd``) builds the targets on the fly and
reduces both BCE terms, and ``hupr_keypoints_argmax`` decodes the keypoints — the only D2H traffic is the ``[B,14,2]`` results.
The loss tensors carry no autograd graph: training differentiates the same loss inside ``hupr_b200.training.TrainStep``
(``hupr_heatmap_loss_bwd``), or through ``HuPRNet.forward`` in ``train()`` mode, whose outputs are autograd-connected (models/networks.py).
"""
import numpy as np
import torch

from .. import ops


def generateTarget(joints, numKeypoints=14, hSize=64, iSize=256):
    """Reference-compatible helper (utils.py:6-65) for ONE sample, evaluated on the GPU: returns (target [K,H,W], kpts [K,2]) numpy."""
    if (numKeypoints, hSize, iSize) != (14, 64, 256):
        raise ValueError("hupr_b200 target synthesis is specialised for 14 keypoints, 64x64 heatmaps, 256-px frames")
    j = torch.as_tensor(np.asarray(joints)).reshape(1, 14, 2)
    dummy = torch.full((1, 14, 64, 64), 0.5, device="cuda")
    _, gt2d, targets = ops.heatmap_loss_fwd(dummy, dummy, j, want_targets=True)
    mu = np.trunc(np.asarray(joints, dtype=np.float64).reshape(14, 2) / (iSize / hSize) + 0.5)
    valid = ~((mu[:, 0] - 6 >= hSize) | (mu[:, 1] - 6 >= hSize) | (mu[:, 0] + 7 < 0) | (mu[:, 1] + 7 < 0))
    return targets[0].cpu().numpy(), mu * valid[:, None]


def get_max_preds(batch_heatmaps):
    """metrics.py:10-38 on the GPU.  Accepts a numpy array or CUDA tensor ``[B, K, 64, 64]``; returns (preds, maxvals) numpy."""
    t = torch.as_tensor(batch_heatmaps)
    if t.dim() != 4:
        raise AssertionError("batch_images should be 4-ndim")
    t = t.to(device="cuda", dtype=torch.float32).contiguous()
    maxvals = torch.empty(t.shape[:2], dtype=torch.float32, device=t.device)
    preds = ops.keypoints_argmax(t, maxvals=maxvals)
    return preds.cpu().numpy(), maxvals.unsqueeze(-1).cpu().numpy()


class _HeatmapBCE(torch.autograd.Function):
    """loss = w1*BCE(heat, T) + w2*BCE(gcn, T) and loss2 = BCE(gcn, T) as autograd nodes over the library kernels: forward =
    hupr_heatmap_loss_fwd (targets synthesised on the fly), backward = the closed form of nn.BCELoss's gradient, (p - T) / (p (1 - p)) / N
    with torch's 1e-12 clamp — so ``loss.backward()`` of the reference loop (tools/run.py:77-78) reaches HuPRNet's train-mode outputs."""

    @staticmethod
    def forward(ctx, heat, gcn, joints, w1, w2):
        losses, gt2d, targets = ops.heatmap_loss_fwd(heat.detach(), gcn.detach(), joints, want_targets=True)
        ctx.save_for_backward(heat.detach(), gcn.detach(), targets)
        ctx.w = (float(w1), float(w2))
        ctx.mark_non_differentiable(gt2d)
        return w1 * losses[2] + w2 * losses[1], losses[1].clone(), gt2d

    @staticmethod
    def backward(ctx, g_loss, g_loss2, _g_gt2d):
        heat, gcn, targets = ctx.saved_tensors
        w1, w2 = ctx.w
        inv_n = 1.0 / heat.numel()
        dh = (heat - targets) / (heat * (1.0 - heat)).clamp_min(1e-12) * (g_loss * (w1 * inv_n))
        dg = (gcn - targets) / (gcn * (1.0 - gcn)).clamp_min(1e-12) * ((g_loss * w2 + g_loss2) * inv_n)
        return dh, dg, None, None, None


class LossComputer(object):
    def __init__(self, cfg, device):
        self.device = device
        self.cfg = cfg
        self.numFrames = cfg.DATASET.numFrames
        self.numGroupFrames = cfg.DATASET.numGroupFrames
        self.numKeypoints = cfg.DATASET.numKeypoints
        self.heatmapSize = self.width = self.height = cfg.DATASET.heatmapSize
        self.imgSize = self.imgWidth = self.imgHeight = cfg.DATASET.imgSize
        self.lossDecay = cfg.TRAINING.lossDecay
        self.alpha = 0.0
        self.beta = 1.0
        if (self.numKeypoints, self.heatmapSize, self.imgSize) != (14, 64, 256):
            raise ValueError("hupr_b200 loss kernels are specialised for 14 keypoints, 64x64 heatmaps, 256-px frames")

    def computeLoss(self, preds, gt, host=True):
        """losses.py:23-45.  ``host=False`` returns pred2d / gt2d as CUDA tensors (no synchronising read-back; the evaluation loop uses it)."""
        heat, gcn = preds
        b = gt.size(0)
        heat = heat.reshape(b, self.numKeypoints, self.height, self.width).contiguous()
        gcn = gcn.reshape(b, self.numKeypoints, self.height, self.width).contiguous()
        if self.alpha < 1.0:
            self.alpha += self.lossDecay
            self.beta -= self.lossDecay
        w1, w2 = (self.alpha, self.beta) if self.lossDecay != -1 else (1.0, 1.0)
        if torch.is_grad_enabled() and (heat.requires_grad or gcn.requires_grad):
            # train-mode predictions of HuPRNet.forward carry an autograd graph: keep the loss connected to it
            loss, loss2, gt2d = _HeatmapBCE.apply(heat, gcn, gt, w1, w2)
        else:
            losses, gt2d, _ = ops.heatmap_loss_fwd(heat, gcn, gt)
            loss2 = losses[1]
            loss = w1 * losses[2] + w2 * losses[1] if self.lossDecay != -1 else losses[0]
        pred2d = ops.keypoints_argmax(gcn.detach())
        if not host:
            return loss, loss2, pred2d, gt2d
        return loss, loss2, pred2d.cpu().numpy(), gt2d.cpu().numpy()
