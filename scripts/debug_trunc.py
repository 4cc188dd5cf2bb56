"""Does the TMEM accumulator truncate?  All-positive operands: a truncating accumulator biases the result low."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from scouter_b200 import _lib as L
dev = torch.device("cuda", 0)
r = np.random.RandomState(0)
Bn, H, W, Cin, Cout, k, g = 4, 14, 14, 256, 128, 3, 1
x = torch.from_numpy(np.abs(r.standard_normal((Bn, H, W, Cin))).astype(np.float32)).to(dev)
w = torch.from_numpy(np.abs(r.standard_normal((Cout, k, k, Cin)) / 48).astype(np.float32)).to(dev)
b = torch.zeros(Cout, device=dev)
ref = torch.nn.functional.conv2d(x.cpu().double().permute(0, 3, 1, 2), w.cpu().double().permute(0, 3, 1, 2), None, 1, 1).permute(0, 2, 3, 1)
for math in (0, 1, 2):
    out = torch.empty(Bn, H, W, Cout, device=dev)
    op = L.Op(kind=L.OP_CONV, src=0, src2=-1, dst=1, cin=Cin, cout=Cout, kh=k, kw=k, stride=1, pad=1, groups=g, flags=0, mid=0,
              reserved=0, w=w.data_ptr(), b=b.data_ptr(), w2=0, b2=0)
    L.check(L.lib().scouter_conv_forward(C.byref(op), x.data_ptr(), 0, out.data_ptr(), Bn, H, W, math, 0))
    torch.cuda.synchronize()
    rel = (out.cpu().double() - ref) / ref
    print(f"chunk={os.environ.get('SCOUTER_UMMA_CHUNK','4')} math={math} K={Cin*9}: signed mean rel err {float(rel.mean()):+.3e}  max |rel| {float(rel.abs().max()):.3e}")
