"""One conv geometry through scouter_conv_forward (MATH_TC, pre-split weights), a few launches -- the target of single-kernel ncu
captures:  ncu --set full -k regex:conv -s 2 -c 1 python scripts/conv_case.py --hw 7 --cin 1024 --cout 2048 --k 1"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scouter_b200 import _lib as L  # noqa: E402
from scouter_b200.plan import split_weights_f16  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--hw", type=int, default=7)
ap.add_argument("--cin", type=int, default=1024)
ap.add_argument("--cout", type=int, default=2048)
ap.add_argument("--k", type=int, default=1)
ap.add_argument("--groups", type=int, default=1)
ap.add_argument("--res", action="store_true")
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, H, k, g = a.batch, a.hw, a.k, a.groups
x = torch.randn(B, H, H, a.cin, device=dev)
w = torch.randn(a.cout, k, k, a.cin // g, device=dev) * (2.0 / (a.cin // g * k * k)) ** 0.5
w2 = split_weights_f16(w)
bias = torch.randn(a.cout, device=dev) * 0.1
res = torch.randn(B, H, H, a.cout, device=dev) if a.res else None
out = torch.empty(B, H, H, a.cout, device=dev)
op = L.Op(kind=L.OP_CONV, src=0, src2=-1, dst=1, cin=a.cin, cout=a.cout, kh=k, kw=k, stride=1, pad=k // 2, groups=g,
          flags=L.F_RELU | (L.F_RESIDUAL if a.res else 0), mid=0, reserved=0, w=w.data_ptr(), b=bias.data_ptr(), w2=w2.data_ptr(), b2=0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for i in range(a.iters):
    if i == a.iters - 1:
        ev[0].record()
    L.check(L.lib().scouter_conv_forward(C.byref(op), x.data_ptr(), L.ptr(res), out.data_ptr(), B, H, H, L.MATH_TC, 0))
ev[1].record()
torch.cuda.synchronize()
print(f"{H}x{H} {a.cin}->{a.cout} k{k} g{g}: {ev[0].elapsed_time(ev[1]) * 1e3:.1f} us")
