"""Signed bias of the tcgen05 conv path (TMEM accumulators truncate on every MMA): mean and max of (out - ref) / |ref| for
all-positive products, all-negative products and mixed signs, 3x3 conv Cin=256 (K = 2304) and 1x1 conv Cin=2048, at the
accumulation chunk given by SCOUTER_UMMA_CHUNK (default 4 k-blocks).   python scripts/probes/accum_bias_probe.py"""
import ctypes as C
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from scouter_b200 import _lib as L  # noqa: E402
from scouter_b200.plan import split_weights_f16  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
print("chunk =", os.environ.get("SCOUTER_UMMA_CHUNK", "4 (default)"))
for name, k, cin, cout, hw in (("3x3 K=2304", 3, 256, 128, 28), ("1x1 K=2048", 1, 2048, 128, 14)):
    for sign in ("++", "+-", "mixed"):
        x = torch.randn(4, hw, hw, cin, device=dev)
        w = torch.randn(cout, k, k, cin, device=dev) * (2.0 / (cin * k * k)) ** 0.5
        if sign != "mixed":
            x = x.abs()
            w = w.abs() if sign == "++" else -w.abs()
        ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), None, 1, k // 2).permute(0, 2, 3, 1)
        w2 = split_weights_f16(w)
        out = torch.empty(4, hw, hw, cout, device=dev)
        op = L.Op(kind=L.OP_CONV, src=0, src2=-1, dst=1, cin=cin, cout=cout, kh=k, kw=k, stride=1, pad=k // 2, groups=1, flags=0, mid=0,
                  reserved=0, w=w.data_ptr(), b=0, w2=w2.data_ptr(), b2=0)
        L.check(L.lib().scouter_conv_forward(C.byref(op), x.data_ptr(), 0, out.data_ptr(), 4, hw, hw, L.MATH_TC, 0))
        torch.cuda.synchronize()
        scale = ref.abs().max()
        e = (out.double() - ref)
        rel = e / ref.abs().clamp_min(1e-3 * float(scale))
        fp32 = F.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), None, 1, k // 2).permute(0, 2, 3, 1)
        e32 = (fp32.double() - ref) / ref.abs().clamp_min(1e-3 * float(scale))
        print(f"{name} {sign:5s}: mean signed rel err {float(rel.mean()):+.2e}, max |err|/max|ref| {float(e.abs().max() / scale):.2e}"
              f"   (torch fp32 conv on the same GPU: mean {float(e32.mean()):+.2e})")
