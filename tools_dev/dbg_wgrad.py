import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from hupr_b200 import ops
from hupr_b200.ops import SplitTensor
n, cin, cout, d, h, w, kernel, pad = 2, 64, 128, 4, 32, 32, (3, 3, 3), (1, 1, 1)
torch.manual_seed(14)
x = torch.randn(n, cin, d, h, w, device="cuda", dtype=torch.float64)
wt = torch.randn(cout, cin, *kernel, device="cuda", dtype=torch.float64, requires_grad=True)
y = F.conv3d(x, wt, padding=pad); dy = torch.randn_like(y); y.backward(dy)
X = SplitTensor.from_float(x.float().permute(0, 2, 3, 4, 1).contiguous())
DY = SplitTensor.from_float(dy.float().permute(0, 2, 3, 4, 1).contiguous())
geom = ops.KMajorGeometry(n, d, h, w, pad)
xt = SplitTensor.empty((cin, geom.ppad), "cuda", zero=True)
dyt = SplitTensor.empty((cout, geom.ppad), "cuda", zero=True)
ops.to_kmajor(X, 0, cin, geom, xt); ops.to_kmajor(DY, 0, cout, geom, dyt)
torch.cuda.synchronize(); print("kmajor ok", geom.ppad)
# check transposed content
P = ((0 * geom.dp + 1 + 1) * geom.hp + 3 + 1) * geom.wp + 5 + 1
print(float(xt.float()[7, P]), float(X.float()[0, 1, 3, 5, 7]))
out = torch.zeros(27, cout, cin, device="cuda")
ops.conv_wgrad(xt, cin, dyt, cout, geom, kernel, out)
torch.cuda.synchronize(); print("wgrad ok")
got = out.permute(1, 2, 0).reshape(cout, cin, *kernel).double()
print(float((got - wt.grad).abs().max() / wt.grad.abs().max()))
