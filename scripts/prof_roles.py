"""Role-level stall accounting of the tcgen05 conv kernels (debug tool, not a bench number).

Builds a second library with -DSCOUTER_PROF (clock64 counters around every mbarrier wait of every warp role), runs
single conv geometries of the B=256 resnest26d forward through scouter_conv_forward and prints, per role, the
average clocks per CTA spent waiting on each barrier.  `python scripts/prof_roles.py [--build-only]`."""
import argparse
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scouter_b200 import _lib as L  # noqa: E402

PROF_LIB = os.path.join(ROOT, "scouter_b200", "libscouter_b200_prof.so")
ap = argparse.ArgumentParser()
ap.add_argument("--build-only", action="store_true")
ap.add_argument("--batch", type=int, default=256)
a = ap.parse_args()

srcs = [os.path.join(L.CSRC, f) for f in os.listdir(L.CSRC)]
if not os.path.exists(PROF_LIB) or any(os.path.getmtime(s) > os.path.getmtime(PROF_LIB) for s in srcs):
    r = subprocess.run(L.nvcc_command(PROF_LIB, ("-DSCOUTER_PROF",)), capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stderr[-3000:])
    print("built", PROF_LIB)
if a.build_only:
    sys.exit(0)

import torch  # noqa: E402
from scouter_b200.plan import split_weights_f16  # noqa: E402

L.LIB_PATH = PROF_LIB
lib = L.lib()
dev = torch.device("cuda", 0)
SLOTS = {0: "prod.total", 1: "prod.wait_patch_empty|empty", 2: "prod.wait_b_empty", 4: "issuer.total", 5: "issuer.wait_pready",
         6: "issuer.wait_cempty", 7: "issuer.wait_bfull|full", 8: "issuer.ksteps", 10: "epi.total", 11: "epi.wait_cfull",
         12: "epi.store", 13: "epi.merge", 14: "split.total", 15: "split.wait_full", 16: "emit.bulk_wait_read", 17: "emit.bar1", 18: "emit.bias+math+sts",
         19: "emit.fence_proxy", 20: "emit.bar2", 21: "emit.tma_issue"}

# (name, H, W, Cin, Cout, k, groups, residual)
CASES = [
    ("conv1.3 halo32", 112, 112, 32, 32, 3, 1, False),
    ("conv1.6 halo64", 112, 112, 32, 64, 3, 1, False),
    ("layer1 conv2 halo64", 56, 56, 64, 128, 3, 2, False),
    ("layer2.0 conv2 halo128", 56, 56, 128, 256, 3, 2, False),
    ("layer3.0 conv2 halo128", 28, 28, 256, 512, 3, 2, False),
    ("layer1 conv3+res flat128", 56, 56, 64, 256, 1, 1, True),
    ("layer1.1 conv1 flat64", 56, 56, 256, 64, 1, 1, False),
    ("layer2 conv3+res flat128", 28, 28, 128, 512, 1, 1, True),
    ("layer2.0 down flat128", 28, 28, 256, 512, 1, 1, False),
    ("layer4 conv3+res flat128", 7, 7, 512, 2048, 1, 1, True),
    ("layer4.0 down flat128", 7, 7, 1024, 2048, 1, 1, False),
]
B = a.batch
for name, H, W, Cin, Cout, k, g, use_res in CASES:
    x = torch.randn(B, H, W, Cin, device=dev)
    w = torch.randn(Cout, k, k, Cin // g, device=dev) * (2.0 / (Cin // g * k * k)) ** 0.5
    w2 = split_weights_f16(w)
    bias = torch.randn(Cout, device=dev) * 0.1
    res = torch.randn(B, H, W, Cout, device=dev) if use_res else None
    out = torch.empty(B, H, W, Cout, device=dev)
    op = L.Op(kind=L.OP_CONV, src=0, src2=-1, dst=1, cin=Cin, cout=Cout, kh=k, kw=k, stride=1, pad=k // 2, groups=g,
              flags=L.F_RELU | (L.F_RESIDUAL if use_res else 0), mid=0, reserved=0, w=w.data_ptr(), b=bias.data_ptr(),
              w2=w2.data_ptr(), b2=0)
    path = lib.scouter_conv_path(C.byref(op), B, H, W, L.MATH_TC)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(3):
        if i == 2:
            ev[0].record()
        L.check(lib.scouter_conv_forward(C.byref(op), x.data_ptr(), L.ptr(res), out.data_ptr(), B, H, W, L.MATH_TC, 0))
    ev[1].record()
    torch.cuda.synchronize()
    host = (C.c_ulonglong * (148 * 32))()
    fn = lib.scouter_prof_read_halo if path == 2 else lib.scouter_prof_read_flat
    fn.argtypes = [C.c_void_p, C.c_int]
    assert fn(host, 148 * 32) == 0
    v = np.array(host[:], dtype=np.float64).reshape(148, 32)
    m = v.mean(0)
    print(f"== {name}: path {path}, {ev[0].elapsed_time(ev[1]) * 1e3:.0f} us; mean clk per CTA:")
    ks = max(m[8], 1.0)
    for s, label in SLOTS.items():
        if m[s]:
            print(f"   {label:30s} {m[s]:12.0f}   per k-step {m[s] / ks:8.1f}")
