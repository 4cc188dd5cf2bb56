"""One eager SlotModel forward inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scouter_b200 as sb
from scouter_b200.synth import make_args
from scouter_b200 import _lib as L
from scouter_b200.synth import fill_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--size", type=int, default=224)
ap.add_argument("--math", default="tc")
ap.add_argument("--classes", type=int, default=10)
ap.add_argument("--spc", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
m = sb.SlotModel(make_args(model="resnest26d", num_classes=a.classes, slots_per_class=a.spc, power=2, to_k_layer=3,
                           loss_status=-1, channel=2048))
m.load_state_dict(fill_state_dict(m.state_dict(), seed=0))
m = m.to(dev).eval()
m.math = {"tc": L.MATH_TC, "fp32": L.MATH_FP32, "tc_fast": L.MATH_TC_FAST}[a.math]
x = torch.randn(a.batch, 3, a.size, a.size, device=dev)
with torch.no_grad():
    for _ in range(2):
        m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one forward, batch", a.batch)
