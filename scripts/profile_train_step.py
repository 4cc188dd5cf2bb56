"""One training step (forward + backward) of SlotModel at a given batch for an ncu launch list (debug tool):
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv python scripts/profile_train_step.py"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scouter_b200 as sb  # noqa: E402
from scouter_b200.synth import fill_state_dict, make_args  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=224)
a = ap.parse_args()
dev = torch.device("cuda", 0)
m = sb.SlotModel(make_args(model="resnest26d", num_classes=10, slots_per_class=1, power=2, to_k_layer=3, loss_status=-1, channel=2048))
m.load_state_dict(fill_state_dict(m.state_dict(), seed=0))
m = m.to(dev).train()
x = torch.randn(a.batch, 3, a.size, a.size, device=dev)
y = torch.randint(0, 10, (a.batch,), device=dev)
for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    out, (loss, nll, attn) = m(x, y)
    loss.backward()
    m.zero_grad(set_to_none=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
