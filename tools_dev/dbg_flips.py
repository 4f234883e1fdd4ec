import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from hupr_b200 import training as TR
from hupr_b200.ops import SplitTensor
from oracle import model as om
dev = "cuda"
def cl(x): return SplitTensor.from_float(x.float().permute(0, 2, 3, 4, 1).contiguous())
def nc(t, c=None):
    v = t.float().permute(0, 4, 1, 2, 3).double()
    return v if c is None else v[:, :c]
def rel(a, b): return float((a - b).norm() / b.norm())
for seed in (0, 1, 2, 3):
    torch.manual_seed(seed)
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
    zc_ref = F.conv2d(x.detach(), sd[prefix + ".main.0.weight"].detach(), padding=1)
    zc = nc(blk.zc)[:, :cout, 0]
    flips = int(((zc > 0) != (zc_ref > 0)).sum())
    s_ref = (F.conv2d(F.prelu(zc_ref, sd[prefix + ".main.1.weight"].detach()), sd[prefix + ".main.2.weight"].detach(), padding=1)
             + F.conv2d(x.detach(), sd[prefix + ".downsample.0.weight"].detach(), padding=1))
    flips2 = int(((nc(blk.s)[:, :cout, 0] > 0) != (s_ref > 0)).sum())
    print("seed", seed, "zc sign flips", flips, "s sign flips", flips2, "dx err", rel(nc(dx, cin)[:, :, 0], x.grad),
          "main.0 err", rel(grads[prefix + ".main.0.weight"].double().reshape(cout, cin, 3, 3), sd[prefix + ".main.0.weight"].grad),
          "main.2 err", rel(grads[prefix + ".main.2.weight"].double().reshape(cout, cout, 3, 3), sd[prefix + ".main.2.weight"].grad))
