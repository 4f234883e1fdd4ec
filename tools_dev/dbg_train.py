import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hupr_b200.models import HuPRNet
from hupr_b200.training import TrainStep
from oracle import model as om
from tests.test_model_gpu import make_cfg
seed, batch = 1, int(sys.argv[1]) if len(sys.argv) > 1 else 1
sd = om.make_state_dict(seed)
hori, vert = om.make_vrdae(batch, seed)
joints = torch.randint(0, 256, (batch, 14, 2), generator=torch.Generator().manual_seed(3))
ref_loss, ref_loss2, ref_grads = om.training_gradients(sd, hori, vert, joints.numpy())
net = HuPRNet(make_cfg()); net.load_state_dict(sd); net = net.cuda().train()
step = TrainStep(net)
loss, loss2 = step.forward_backward(hori.cuda(), vert.cuda(), joints)
torch.cuda.synchronize()
print("loss", float(loss), ref_loss, float(loss2), ref_loss2)
for name, q in net.named_parameters():
    ref = ref_grads[name]
    err = float((q.grad.cpu() - ref).norm() / ref.norm().clamp_min(1e-30))
    print("%-60s %.3g   ref max %.3g" % (name, err, float(ref.abs().max())))
