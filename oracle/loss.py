"""Oracle (TEST INFRASTRUCTURE): numpy / torch-CPU restatement of the heatmap-loss tail of the hot path.

Follows /root/reference/misc:
  generateTarget   utils.py:6-65     13x13 un-normalised Gaussian (sigma 2) pasted at mu = int(j / 4 + 0.5), clipped to 64x64
  get_max_preds    metrics.py:10-38  flat argmax -> (x = idx % W, y = floor(idx / W)), zeroed where max <= 0
  computeLoss      losses.py:23-45   loss1 = BCE(heatmap, T), loss2 = BCE(gcn_heatmap, T), loss = loss1 + loss2 (lossDecay == -1)

Pinned against the reference functions by tests/golden/loss_reference.npz (oracle/make_golden.py loss).
"""
import numpy as np
import torch
import torch.nn.functional as F

HEATMAP = 64
IMG = 256
SIGMA = 2
NUM_KEYPOINTS = 14


def generate_target(joints, num_keypoints=NUM_KEYPOINTS, hsize=HEATMAP, isize=IMG):
    """joints: int array [K, 2] in image pixels -> (target float32 [K, H, W], target_kpts float64 [K, 2])."""
    joints = np.asarray(joints)
    target = np.zeros((num_keypoints, hsize, hsize), dtype=np.float32)
    kpts = np.zeros((num_keypoints, 2))
    tmp = SIGMA * 3
    stride = isize / hsize
    ax = np.arange(0, 2 * tmp + 1, 1, np.float32)
    g = np.exp(-((ax[None, :] - tmp) ** 2 + (ax[:, None] - tmp) ** 2) / (2 * SIGMA ** 2))   # float32 like the reference
    for k in range(num_keypoints):
        mu_x = int(joints[k][0] / stride + 0.5)
        mu_y = int(joints[k][1] / stride + 0.5)
        ul = (mu_x - tmp, mu_y - tmp)
        br = (mu_x + tmp + 1, mu_y + tmp + 1)
        if ul[0] >= hsize or ul[1] >= hsize or br[0] < 0 or br[1] < 0:
            continue
        gx = (max(0, -ul[0]), min(br[0], hsize) - ul[0])
        gy = (max(0, -ul[1]), min(br[1], hsize) - ul[1])
        ix = (max(0, ul[0]), min(br[0], hsize))
        iy = (max(0, ul[1]), min(br[1], hsize))
        target[k, iy[0]:iy[1], ix[0]:ix[1]] = g[gy[0]:gy[1], gx[0]:gx[1]]
        kpts[k] = (mu_x, mu_y)
    return target, kpts


def get_max_preds(heatmaps):
    """heatmaps [B, K, H, W] -> (preds float32 [B, K, 2] = (x, y), maxvals [B, K, 1])."""
    heatmaps = np.asarray(heatmaps)
    b, k, h, w = heatmaps.shape
    flat = heatmaps.reshape(b, k, -1)
    idx = flat.argmax(2)
    maxvals = flat.max(2)[..., None]
    preds = np.stack([idx % w, idx // w], axis=2).astype(np.float32)
    preds *= (maxvals > 0.0).astype(np.float32)
    return preds, maxvals


def compute_loss(heatmap, gcn_heatmap, gt):
    """heatmap [B,14,1,64,64], gcn_heatmap [B,1,14,64,64] (torch fp32, post-sigmoid), gt int [B,14,2].
    Returns (loss, loss2, pred2d [B,14,2], gt2d [B,14,2], targets [B,14,64,64])."""
    gt = np.asarray(gt)
    b = gt.shape[0]
    targets = np.stack([generate_target(gt[i])[0] for i in range(b)])
    t = torch.from_numpy(targets)
    loss1 = F.binary_cross_entropy(heatmap.reshape(b, NUM_KEYPOINTS, HEATMAP, HEATMAP), t)
    loss2 = F.binary_cross_entropy(gcn_heatmap.reshape(b, NUM_KEYPOINTS, HEATMAP, HEATMAP), t)
    pred2d, _ = get_max_preds(gcn_heatmap.reshape(b, NUM_KEYPOINTS, HEATMAP, HEATMAP).numpy())
    gt2d, _ = get_max_preds(targets)
    return float(loss1 + loss2), float(loss2), pred2d, gt2d, targets
