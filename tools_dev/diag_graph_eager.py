import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hupr_b200.models import HuPRNet
from hupr_b200.training import TrainStep
from oracle import model as om
from tests.test_model_gpu import make_cfg
sd = om.make_state_dict(4)
hori, vert = (t.cuda() for t in om.make_vrdae(1, 4))
joints = torch.randint(0, 256, (1, 14, 2), generator=torch.Generator().manual_seed(5)).cuda()
for mode in ("eager", "eager", "graph", "graph"):
    net = HuPRNet(make_cfg()); net.load_state_dict(sd); net = net.cuda().train()
    step = TrainStep(net)
    losses = []
    if mode == "eager":
        for _ in range(5):
            loss, _ = step.forward_backward(hori, vert, joints); step.optimizer_step(); losses.append(float(loss))
    else:
        replay = step.capture(hori, vert, joints)
        for _ in range(3):
            loss, _ = replay(); losses.append(float(loss))
        losses = [None, None] + losses
    print(mode, losses, flush=True)
# gradient determinism of one forward_backward
net = HuPRNet(make_cfg()); net.load_state_dict(sd); net = net.cuda().train()
step = TrainStep(net)
step.forward_backward(hori, vert, joints); g1 = step.flat_g.clone()
step.forward_backward(hori, vert, joints); g2 = step.flat_g.clone()
if g1 is not None:
    print("grad rel L2 diff between two identical eager passes:", float((g1 - g2).norm() / g1.norm()), "max abs", float((g1 - g2).abs().max()), float(g1.abs().max()))
