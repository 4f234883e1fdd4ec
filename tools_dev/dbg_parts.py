import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from hupr_b200 import ops, training as TR
from hupr_b200.ops import SplitTensor
from oracle import model as om
torch.manual_seed(0)
dev = "cuda"
def cl(x): return SplitTensor.from_float(x.float().permute(0, 2, 3, 4, 1).contiguous())
def nc(t, c=None):
    v = t.float().permute(0, 4, 1, 2, 3).double()
    return v if c is None else v[:, :c]
def rel(a, b): return float((a - b).norm() / b.norm())      # relative L2: isolated ReLU-mask flips (|pre-activation| < 1e-5) do not dominate

# ---------------- Block2D
cin, cout, hw, b = 128, 64, 32, 2
prefix = "blk"
sd = {prefix + ".main.0.weight": torch.randn(cout, cin, 3, 3, device=dev, dtype=torch.float64) * 0.05,
      prefix + ".main.1.weight": torch.tensor([0.3], device=dev, dtype=torch.float64),
      prefix + ".main.2.weight": torch.randn(cout, cout, 3, 3, device=dev, dtype=torch.float64) * 0.05,
      prefix + ".downsample.0.weight": torch.randn(cout, cin, 3, 3, device=dev, dtype=torch.float64) * 0.05,
      prefix + ".relu.weight": torch.tensor([0.2], device=dev, dtype=torch.float64)}
for v in sd.values(): v.requires_grad_()
x = torch.randn(b, cin, hw, hw, device=dev, dtype=torch.float64, requires_grad=True)
y = om.block2d(x, sd, prefix); gy = torch.randn_like(y); y.backward(gy)
blk = TR.Block2D(prefix, cin, cout); blk.pack({k: v.detach().float() for k, v in sd.items()})
out = blk.forward(cl(x.detach().unsqueeze(2)))
grads = {}
dx = blk.backward(cl(gy.unsqueeze(2)), 0, grads)
torch.cuda.synchronize()
print("Block2D fwd", rel(nc(out, cout)[:, :, 0], y.detach()), "dx", rel(nc(dx, cin)[:, :, 0], x.grad))
for k in sd: print("   ", k, rel(grads[k].double().reshape(sd[k].shape), sd[k].grad))

# ---------------- Block3D (train)
cin, cout, d, hw, b = 64, 128, 4, 32, 1
prefix = "b3"
sd = {}
for n, shp in ((".main.0.weight", (cout, cin, 3, 3, 3)), (".main.3.weight", (cout, cout, 3, 3, 3)), (".downsample.0.weight", (cout, cin, 3, 3, 3))):
    sd[prefix + n] = torch.randn(shp, device=dev, dtype=torch.float64) * 0.03
for bn in (".main.1", ".main.4", ".downsample.1"):
    sd[prefix + bn + ".weight"] = torch.rand(cout, device=dev, dtype=torch.float64) + 0.5
    sd[prefix + bn + ".bias"] = torch.randn(cout, device=dev, dtype=torch.float64) * 0.1
    sd[prefix + bn + ".running_mean"] = torch.zeros(cout, device=dev, dtype=torch.float64)
    sd[prefix + bn + ".running_var"] = torch.ones(cout, device=dev, dtype=torch.float64)
for k, v in sd.items():
    if "running" not in k: v.requires_grad_()
x = torch.randn(b, cin, d, hw, hw, device=dev, dtype=torch.float64, requires_grad=True)
om.TRAINING = True
y = om.block3d(x, sd, prefix); gy = torch.randn_like(y); y.backward(gy)
om.TRAINING = False
blk = TR.Block3D(prefix, cin, cout)
params = {k: v.detach().float() for k, v in sd.items() if "running" not in k}
buffers = {k: v.detach().float().clone() for k, v in sd.items() if "running" in k}
for bn in (".main.1", ".main.4", ".downsample.1"): buffers[prefix + bn + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long, device=dev)
blk.pack(params)
out = blk.forward(cl(x.detach()), params, buffers)
grads = {}
dx = blk.backward(cl(gy), grads)
torch.cuda.synchronize()
print("Block3D fwd", rel(nc(out), y.detach()), "dx", rel(nc(dx, cin), x.grad))
for k in params: print("   ", k, rel(grads[k].double().reshape(sd[k].shape), sd[k].grad))

# ---------------- AttentionLevel
for c, hw in ((64, 16), (256, 16), (128, 32)):
    b, prev = 2, 64
    s = hw * hw
    lvl = TR.AttentionLevel(0, c, hw, prev)
    sd = {}
    for n in TR.L.PROJ_HORI + TR.L.PROJ_VERT:
        sd["radarDecoder.%s.0.weight" % n] = (torch.randn(c, c, 1, 1, device=dev, dtype=torch.float64) / c ** 0.5 * 0.7).requires_grad_()
    ra = torch.randn(b, c, hw, hw, device=dev, dtype=torch.float64, requires_grad=True)
    re = torch.randn(b, c, hw, hw, device=dev, dtype=torch.float64, requires_grad=True)
    outs = om.attention_level(ra, re, sd, 0)
    y = torch.cat(outs, 1); gy = torch.randn_like(y); y.backward(gy)
    lvl.pack({k: v.detach().float() for k, v in sd.items()})
    cat = SplitTensor.empty((b, 1, hw, hw, prev + 4 * c), dev, zero=True)
    lvl.forward(cl(ra.detach().unsqueeze(2)), cl(re.detach().unsqueeze(2)), cat)
    dcat = torch.zeros(b, prev + 4 * c, 1, hw, hw, device=dev, dtype=torch.float64); dcat[:, prev:] = gy.unsqueeze(2)
    grads = {}
    dra, dre = lvl.backward(cl(dcat), grads)
    torch.cuda.synchronize()
    print("Attention c=%d S=%d fwd" % (c, s), rel(nc(cat)[:, prev:, 0], y.detach()), "dra", rel(nc(dra)[:, :, 0], ra.grad), "dre", rel(nc(dre)[:, :, 0], re.grad))
    for k in sd: print("   ", k, rel(grads[k].double().reshape(sd[k].shape), sd[k].grad))
