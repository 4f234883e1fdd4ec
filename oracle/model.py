"""Oracle (TEST INFRASTRUCTURE, never on the product path): torch-CPU fp32 functional restatement of
MSCSA-PRGCN (HuPRNet) inference, driven by a reference-format ``state_dict`` (255 entries).

Follows /root/reference/models:
  forward_chirp / MNet          networks.py:23-33, chirp_networks.py:11-21
  Encoder3D / BasicBlock3D      layers.py:186-217, :40-70
  attention + decoder           layers.py:126-184, BasicBlock2D :8-38
  PRGCN / GCN_layers            gcn_networks.py:6-64, adjacency layers.py:97-112
  HuPRNet.forward               networks.py:35-41

Pinned against the reference module itself (same state_dict, same input) by tests/golden/model_*.npz —
see oracle/make_golden.py.  ``make_state_dict(seed)`` builds the seeded synthetic weights used by the tests,
the smoke test and the bench (no checkpoint is available offline).
"""
import math

import torch
import torch.nn.functional as F

NUM_FILTERS = 32
NUM_KEYPOINTS = 14
GROUP_FRAMES = 8
CHIRPS = 8

ADJACENCY = torch.tensor([
    [1, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0],
    [0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 1, 0, 1, 1, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0],
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1],
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1],
], dtype=torch.float32)   # layers.py:97-112 (asymmetric, un-normalised)


# ------------------------------------------------------------------------------------------------ weights
def state_dict_spec(nf=NUM_FILTERS):
    """Ordered (name, shape) list of the 255 reference state_dict entries (construction order, networks.py:17-21)."""
    spec = []

    def bn(prefix, c):
        spec.extend([(prefix + ".weight", (c,)), (prefix + ".bias", (c,)), (prefix + ".running_mean", (c,)),
                     (prefix + ".running_var", (c,)), (prefix + ".num_batches_tracked", ())])

    def block3d(prefix, cin, cout):
        spec.append((prefix + ".main.0.weight", (cout, cin, 3, 3, 3)))
        bn(prefix + ".main.1", cout)
        spec.append((prefix + ".main.3.weight", (cout, cout, 3, 3, 3)))
        bn(prefix + ".main.4", cout)
        spec.append((prefix + ".downsample.0.weight", (cout, cin, 3, 3, 3)))
        bn(prefix + ".downsample.1", cout)

    def block2d(prefix, cin, cout):
        spec.append((prefix + ".main.0.weight", (cout, cin, 3, 3)))
        spec.append((prefix + ".main.1.weight", (1,)))
        spec.append((prefix + ".main.2.weight", (cout, cout, 3, 3)))
        spec.append((prefix + ".downsample.0.weight", (cout, cin, 3, 3)))
        spec.append((prefix + ".relu.weight", (1,)))

    for net in ("RAchirpNet", "REchirpNet"):
        spec.append((net + ".temporalConvWx1x1.weight", (nf, 2, 2, 1, 1)))
        spec.append((net + ".temporalConvWx1x1.bias", (nf,)))
    for enc in ("RAradarEncoder", "REradarEncoder"):
        spec.append((enc + ".layer1.0.weight", (nf * 2, nf, 3, 3, 3)))
        spec.append((enc + ".layer1.0.bias", (nf * 2,)))
        block3d(enc + ".layer1.1", nf * 2, nf * 2)
        block3d(enc + ".layer2.1", nf * 2, nf * 4)
        block3d(enc + ".layer2.2", nf * 4, nf * 4)
        block3d(enc + ".layer3.1", nf * 4, nf * 8)
        block3d(enc + ".layer3.2", nf * 8, nf * 8)
        spec.append((enc + ".l1temporalMerge.weight", (nf * 2, nf * 2, GROUP_FRAMES, 1, 1)))
        spec.append((enc + ".l2temporalMerge.weight", (nf * 4, nf * 4, GROUP_FRAMES // 2, 1, 1)))
        spec.append((enc + ".temporalMerge.weight", (nf * 8, nf * 8, GROUP_FRAMES // 4, 1, 1)))
    dec = "radarDecoder"
    block2d(dec + ".decoderLayer3.0", nf * 32, nf * 8)
    block2d(dec + ".decoderLayer3.1", nf * 8, nf * 4)
    block2d(dec + ".decoderLayer2.0", nf * 20, nf * 4)
    block2d(dec + ".decoderLayer2.1", nf * 4, nf * 2)
    block2d(dec + ".decoderLayer1.0", nf * 10, nf * 2)
    block2d(dec + ".decoderLayer1.1", nf * 2, nf)
    spec.append((dec + ".decoderLayer1.2.weight", (NUM_KEYPOINTS, nf, 1, 1)))
    for layer in ("L1", "L2", "L3"):
        spec.append((dec + ".gcn.%s.weight" % layer, (1024, 1024)))
        spec.append((dec + ".gcn.%s.bias" % layer, (1024, NUM_KEYPOINTS)))
    for name in ("phi_cross_hori", "theta_cross_hori", "phi_cross_vert", "theta_cross_vert",
                 "phi_self_hori", "theta_self_hori", "phi_self_vert", "theta_self_vert"):
        for i, c in enumerate((nf * 8, nf * 4, nf * 2)):
            spec.append((dec + ".%s.%d.weight" % (name, i), (c, c, 1, 1)))
    return spec


def make_state_dict(seed=0, nf=NUM_FILTERS):
    """Seeded synthetic weights in the reference's state_dict format.

    Scales follow torch's default inits (uniform +-1/sqrt(fan_in)); BatchNorm affine/running statistics and PReLU
    slopes are randomised so that folding / per-channel epilogues are actually exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in state_dict_spec(nf):
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(100, dtype=torch.long)
        elif name.endswith("running_var"):
            sd[name] = torch.rand(shape, generator=g) * 0.5 + 0.25
        elif name.endswith("running_mean"):
            sd[name] = torch.randn(shape, generator=g) * 0.1
        elif ".main.1.weight" in name and shape == (1,) or name.endswith("relu.weight"):
            sd[name] = torch.rand(shape, generator=g) * 0.3 + 0.1                       # PReLU slope
        elif len(shape) == 1 and (".main.1." in name or ".main.4." in name or ".downsample.1." in name):
            if name.endswith(".weight"):
                sd[name] = torch.rand(shape, generator=g) * 0.5 + 0.75                  # BN gamma
            else:
                sd[name] = torch.randn(shape, generator=g) * 0.1                        # BN beta
        else:
            fan_in = shape[1] if len(shape) == 2 else (int(torch.tensor(shape[1:]).prod()) if len(shape) > 1 else shape[0])
            if name.endswith(".bias") and "gcn" in name:
                fan_in = 1024
            if name.endswith("temporalConvWx1x1.bias"):
                fan_in = 4
            if name.endswith("layer1.0.bias"):
                fan_in = nf * 27
            # LeCun-uniform (var = 1/fan_in) for the conv stacks keeps activations O(1) through 20+ layers (so that relative-error
            # checks on logits are meaningful); GCN / biases keep torch's +-1/sqrt(fan_in)
            gain = math.sqrt(3.0) if (len(shape) >= 4) else 1.0
            bound = gain / math.sqrt(fan_in)
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def make_vrdae(batch, seed=0):
    """Seeded synthetic network inputs ``[B, 8, 8, 2, 64, 64, 8]`` x 2 (standardised planes are ~N(0,1))."""
    g = torch.Generator().manual_seed(1000 + seed)
    shape = (batch, GROUP_FRAMES, CHIRPS, 2, 64, 64, 8)
    return torch.randn(shape, generator=g), torch.randn(shape, generator=g)


# ------------------------------------------------------------------------------------------------ forward
TRAINING = False      # set by huprnet_forward(training=True): BatchNorm uses batch statistics (model.train(), tools/run.py:66)


def _bn(x, sd, prefix):
    if TRAINING:
        return F.batch_norm(x, sd[prefix + ".running_mean"].clone(), sd[prefix + ".running_var"].clone(), sd[prefix + ".weight"],
                            sd[prefix + ".bias"], True, 0.1, 1e-5)
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], False, 0.1, 1e-5)


def chirp_net(vrdae, sd, prefix):
    """networks.py:23-33 for one sensor: elevation mean, the (chirp, re/im)->(channel, depth) .view, MNet."""
    b = vrdae.size(0)
    x = vrdae.mean(dim=6).view(b * GROUP_FRAMES, -1, CHIRPS, 64, 64)
    x = F.conv3d(x, sd[prefix + ".temporalConvWx1x1.weight"], sd[prefix + ".temporalConvWx1x1.bias"], stride=(2, 1, 1))
    x = F.max_pool3d(x, (CHIRPS // 2, 1, 1), (CHIRPS // 2, 1, 1))
    return x.squeeze(2).view(b, GROUP_FRAMES, -1, 64, 64).permute(0, 2, 1, 3, 4)


def block3d(x, sd, prefix):
    res = _bn(F.conv3d(x, sd[prefix + ".downsample.0.weight"], padding=1), sd, prefix + ".downsample.1")
    y = F.relu(_bn(F.conv3d(x, sd[prefix + ".main.0.weight"], padding=1), sd, prefix + ".main.1"))
    y = _bn(F.conv3d(y, sd[prefix + ".main.3.weight"], padding=1), sd, prefix + ".main.4")
    return F.relu(y + res)


def encoder3d(x, sd, prefix):
    l1 = F.conv3d(x, sd[prefix + ".layer1.0.weight"], sd[prefix + ".layer1.0.bias"], padding=1)
    l1 = block3d(l1, sd, prefix + ".layer1.1")
    l2 = F.interpolate(l1, scale_factor=0.5, mode="trilinear", align_corners=True)
    l2 = block3d(block3d(l2, sd, prefix + ".layer2.1"), sd, prefix + ".layer2.2")
    l3 = F.interpolate(l2, scale_factor=0.5, mode="trilinear", align_corners=True)
    l3 = block3d(block3d(l3, sd, prefix + ".layer3.1"), sd, prefix + ".layer3.2")
    return (F.conv3d(l1, sd[prefix + ".l1temporalMerge.weight"]).squeeze(2),
            F.conv3d(l2, sd[prefix + ".l2temporalMerge.weight"]).squeeze(2),
            F.conv3d(l3, sd[prefix + ".temporalMerge.weight"]).squeeze(2))


def attention(k, q, v):
    """layers.py:126-133: logits[key, query] = <k, q>, softmax over the KEY axis, no 1/sqrt(d)."""
    b, c, h, w = v.shape
    logits = torch.einsum("bij,bik->bjk", k.view(b, c, h * w), q.view(b, c, h * w))
    out = torch.einsum("bci,bik->bck", v.view(b, c, h * w), F.softmax(logits, 1))
    return out.view(b, c, h, w)


def block2d(x, sd, prefix):
    res = F.conv2d(x, sd[prefix + ".downsample.0.weight"], padding=1)
    y = F.prelu(F.conv2d(x, sd[prefix + ".main.0.weight"], padding=1), sd[prefix + ".main.1.weight"])
    y = F.conv2d(y, sd[prefix + ".main.2.weight"], padding=1)
    return F.prelu(y + res, sd[prefix + ".relu.weight"])


def attention_level(ra, re, sd, level):
    p = "radarDecoder."
    conv = lambda name, x: F.conv2d(x, sd[p + "%s.%d.weight" % (name, level)])
    ra_cross = attention(conv("phi_cross_hori", ra), conv("theta_cross_vert", re), ra) + ra
    ra_self = attention(conv("phi_self_hori", ra), conv("theta_self_hori", ra), ra)
    re_cross = attention(conv("phi_cross_vert", re), conv("theta_cross_hori", ra), re) + re
    re_self = attention(conv("phi_self_vert", re), conv("theta_self_vert", re), re)
    return [ra_cross, ra_self, re_cross, re_self]


def prgcn(logits, sd):
    """gcn_networks.py:47-64."""
    b = logits.size(0)
    x = F.interpolate(logits, scale_factor=0.5, mode="bilinear", align_corners=True)
    x = x.reshape(-1, NUM_KEYPOINTS, 1024).permute(0, 2, 1)
    for i, layer in enumerate(("L1", "L2", "L3")):
        x = torch.matmul(sd["radarDecoder.gcn.%s.weight" % layer], torch.matmul(x, ADJACENCY)) + sd["radarDecoder.gcn.%s.bias" % layer]
        if i < 2:
            x = F.relu(x)
    x = x.permute(0, 2, 1).reshape(b, NUM_KEYPOINTS, 32, 32)
    x = F.interpolate(x, scale_factor=2.0, mode="bilinear", align_corners=True)
    return torch.sigmoid(x).unsqueeze(1)


def decoder(feats_ra, feats_re, sd):
    """layers.py:135-184.  feats_* = (l1, l2, l3) maps of one sensor."""
    p = "radarDecoder."
    up = lambda x: F.interpolate(x, scale_factor=2.0, mode="bilinear", align_corners=True)
    maps = torch.cat(attention_level(feats_ra[2], feats_re[2], sd, 0), 1)
    maps = up(block2d(block2d(maps, sd, p + "decoderLayer3.0"), sd, p + "decoderLayer3.1"))
    maps = torch.cat([maps] + attention_level(feats_ra[1], feats_re[1], sd, 1), 1)
    maps = up(block2d(block2d(maps, sd, p + "decoderLayer2.0"), sd, p + "decoderLayer2.1"))
    maps = torch.cat([maps] + attention_level(feats_ra[0], feats_re[0], sd, 2), 1)
    maps = block2d(block2d(maps, sd, p + "decoderLayer1.0"), sd, p + "decoderLayer1.1")
    logits = F.conv2d(maps, sd[p + "decoderLayer1.2.weight"])
    return logits, prgcn(logits, sd)


def huprnet_forward(sd, vrdae_hori, vrdae_vert, return_intermediates=False, training=False):
    """HuPRNet.forward (networks.py:35-41): returns (heatmap [B,14,1,64,64], gcn_heatmap [B,1,14,64,64])."""
    global TRAINING
    TRAINING = training
    try:
        return _huprnet_forward(sd, vrdae_hori, vrdae_vert, return_intermediates)
    finally:
        TRAINING = False


def training_gradients(sd, vrdae_hori, vrdae_vert, joints):
    """Reference training step semantics (tools/run.py:74-78): model.train() forward, loss = BCE + BCE, loss.backward().
    Returns (loss, loss2, {parameter name: gradient}) computed by torch autograd on the CPU."""
    from . import loss as oloss
    import numpy as np
    names = [k for k, v in sd.items() if v.dtype.is_floating_point and "running_" not in k]
    leaf = {k: (sd[k].clone().requires_grad_() if k in names else sd[k]) for k in sd}
    heat, gcn = huprnet_forward(leaf, vrdae_hori, vrdae_vert, training=True)
    b = vrdae_hori.shape[0]
    targets = torch.from_numpy(np.stack([oloss.generate_target(np.asarray(joints[i]))[0] for i in range(b)]))
    loss1 = F.binary_cross_entropy(heat.reshape(b, NUM_KEYPOINTS, 64, 64), targets)
    loss2 = F.binary_cross_entropy(gcn.reshape(b, NUM_KEYPOINTS, 64, 64), targets)
    (loss1 + loss2).backward()
    return float((loss1 + loss2).detach()), float(loss2.detach()), {k: leaf[k].grad for k in names}


def _huprnet_forward(sd, vrdae_hori, vrdae_vert, return_intermediates=False):
    ra = chirp_net(vrdae_hori, sd, "RAchirpNet")
    re = chirp_net(vrdae_vert, sd, "REchirpNet")
    feats_ra = encoder3d(ra, sd, "RAradarEncoder")
    feats_re = encoder3d(re, sd, "REradarEncoder")
    logits, gcn_heatmap = decoder(feats_ra, feats_re, sd)
    heatmap = torch.sigmoid(logits).unsqueeze(2)
    if return_intermediates:
        return heatmap, gcn_heatmap, dict(ra=ra, re=re, feats_ra=feats_ra, feats_re=feats_re, logits=logits)
    return heatmap, gcn_heatmap
