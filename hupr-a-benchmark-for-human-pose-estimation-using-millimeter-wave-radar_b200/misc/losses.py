"""Heatmap loss on the B200: host-side mirror of /root/reference/misc/losses.py (``LossComputer``), misc/utils.py
(``generateTarget``) and misc/metrics.py (``get_max_preds``).

``LossComputer(cfg, device).computeLoss(preds, gt) -> (loss, loss2, pred2d, gt2d)`` keeps the reference's signature and return
convention (losses.py:23-45): ``loss``/``loss2`` are 0-dim float32 tensors on the device, ``pred2d``/``gt2d`` are float32 numpy
arrays ``[B, 14, 2]`` of heatmap-pixel (x, y).  The reference synthesises the Gaussian targets on the CPU, copies them to the GPU
and copies the heatmaps back for a numpy argmax; here one kernel pair (``hupr_heatmap_loss_fwd``) builds the targets on the fly and
reduces both BCE terms, and ``hupr_keypoints_argmax`` decodes the keypoints — the only D2H traffic is the ``[B,14,2]`` results.
Forward only: the loss tensors carry no autograd graph (the training backward is not built yet).
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
        losses, gt2d, _ = ops.heatmap_loss_fwd(heat, gcn, gt)
        pred2d = ops.keypoints_argmax(gcn)
        if self.alpha < 1.0:
            self.alpha += self.lossDecay
            self.beta -= self.lossDecay
        if self.lossDecay != -1:
            loss = self.alpha * losses[2] + self.beta * losses[1]
        else:
            loss = losses[0]
        if not host:
            return loss, losses[1], pred2d, gt2d
        return loss, losses[1], pred2d.cpu().numpy(), gt2d.cpu().numpy()
