"""One steady-state, graph-replayed training step between cudaProfilerStart/Stop — for `ncu --profile-from-start off` launch lists
(every kernel of the step, hupr:: or not) and for torch.profiler traces.  argv: batch [products]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_cfg
from hupr_b200.models import HuPRNet
from hupr_b200.training import TrainStep

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
products = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda")
torch.manual_seed(0)
model = HuPRNet(make_cfg()).to(dev).train()
trainer = TrainStep(model, products=products)
gen = torch.Generator(device=dev).manual_seed(1)
hori = torch.randn((b, 8, 8, 2, 64, 64, 8), generator=gen, device=dev)
vert = torch.randn((b, 8, 8, 2, 64, 64, 8), generator=gen, device=dev)
joints = torch.randint(0, 256, (b, 14, 2), generator=gen, device=dev)
replay = trainer.capture(hori, vert, joints)
for _ in range(2):
    replay()
torch.cuda.synchronize()
torch.cuda.profiler.start()
replay()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done; step count", trainer.step_count)
