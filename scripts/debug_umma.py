"""Debug aid: tcgen05 conv vs the exact CUDA-core conv on the device, with an error map per case."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from scouter_b200 import _lib as L
from scouter_b200.plan import round_tf32

dev = torch.device("cuda", 0)
CASES = [(1, 8, 16, 32, 32, 1, 1), (2, 56, 56, 64, 64, 1, 1), (1, 8, 16, 32, 32, 3, 1), (2, 56, 56, 64, 128, 3, 2),
         (4, 7, 7, 2048, 512, 1, 1), (9, 7, 7, 512, 1024, 3, 2), (256, 7, 7, 2048, 128, 1, 1)]
for (Bn, H, W, Cin, Cout, k, g) in CASES:
    r = np.random.RandomState(1)
    x = round_tf32(torch.from_numpy(r.standard_normal((Bn, H, W, Cin)).astype(np.float32))).to(dev)
    w = round_tf32(torch.from_numpy((r.standard_normal((Cout, k, k, Cin // g)) / np.sqrt(Cin // g * k * k)).astype(np.float32))).to(dev)
    b = torch.zeros(Cout, device=dev)
    outs = []
    for math in (0, 1):
        out = torch.full((Bn, H, W, Cout), float("nan"), device=dev)
        op = L.Op(kind=L.OP_CONV, src=0, src2=-1, dst=1, cin=Cin, cout=Cout, kh=k, kw=k, stride=1, pad=k // 2, groups=g,
                  flags=0, mid=0, reserved=0, w=w.data_ptr(), b=b.data_ptr(), w2=0, b2=0)
        rc = L.lib().scouter_conv_forward(C.byref(op), x.data_ptr(), 0, out.data_ptr(), Bn, H, W, math, 0)
        if rc:
            print("rc", rc, L.lib().scouter_last_error())
        torch.cuda.synchronize()
        outs.append(out.cpu())
    a, t = outs
    err = (a - t).abs()
    nan = int(torch.isnan(t).sum())
    print(f"case B{Bn} {H}x{W} c{Cin}->{Cout} k{k} g{g}: max err {float(err.nan_to_num(1e9).max()):.3e} "
          f"ref max {float(a.abs().max()):.3f} nan {nan}/{t.numel()}")
    if float(err.nan_to_num(1e9).max()) > 1e-2:
        e = err.nan_to_num(1e9).view(-1, Cout)
        rows_bad = (e.max(1).values > 1e-2).nonzero().flatten()[:16].tolist()
        cols_bad = (e.max(0).values > 1e-2).nonzero().flatten()[:16].tolist()
        print("   first bad rows", rows_bad, "bad cols", cols_bad)
        print("   ref[0,:8]", a.view(-1, Cout)[0, :8].tolist())
        print("   got[0,:8]", t.view(-1, Cout)[0, :8].tolist())
