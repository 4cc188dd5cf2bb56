"""CPU: logic check of the DRAFT AdamW step (row f1; scouter_b200/csrc/draft/adamw.cuh, not in the library) against
``torch.optim.AdamW(params, lr=...)`` exactly as the reference constructs it (train.py:146), by host emulation."""
import ctypes as C
import math
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

DRAFT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scouter_b200", "csrc", "draft")
_f = C.POINTER(C.c_float)


class Args(C.Structure):            # scouter_draft::AdamWArgs
    _fields_ = [(k, C.c_float) for k in ("decay", "one_minus_beta1", "beta2", "one_minus_beta2", "eps", "step_size",
                                         "bias_correction2_sqrt")] + \
               [("n", C.c_size_t), ("p", _f), ("g", _f), ("m", _f), ("v", _f)]


def test_adamw_draft_matches_torch(tmp_path):
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    so = str(tmp_path / "adamw_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", os.path.join(DRAFT, "adamw_host.cpp"),
                    "-o", so], check=True)
    lib = C.CDLL(so)
    lib.adamw_host.argtypes = [C.POINTER(Args)]
    lib.adamw_host.restype = None
    g = torch.Generator().manual_seed(0)
    n, lr = 10007, 1e-4                                        # the reference's default --lr
    w = torch.nn.Parameter(torch.randn(n, generator=g))
    opt = torch.optim.AdamW([w], lr=lr)                        # train.py:146: defaults for everything else
    d = opt.defaults
    p = w.detach().numpy().copy()
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in range(1, 6):
        grad = torch.randn(n, generator=g) * (10.0 ** (step - 3))
        w.grad = grad.clone()
        opt.step()
        gn = grad.numpy().copy()
        b1, b2 = d["betas"]
        a = Args(decay=1 - lr * d["weight_decay"], one_minus_beta1=1 - b1, beta2=b2, one_minus_beta2=1 - b2, eps=d["eps"],
                 step_size=lr / (1 - b1 ** step), bias_correction2_sqrt=math.sqrt(1 - b2 ** step), n=n,
                 p=p.ctypes.data_as(_f), g=gn.ctypes.data_as(_f), m=m.ctypes.data_as(_f), v=v.ctypes.data_as(_f))
        lib.adamw_host(C.byref(a))
        st = opt.state[w]
        for mine, ref in ((m, st["exp_avg"].numpy()), (v, st["exp_avg_sq"].numpy())):
            assert np.abs(mine - ref).max() <= 1e-6 * np.abs(ref).max(), step        # a few fp32 ulps (measured 1-2e-7)
        assert np.abs(p - w.detach().numpy()).max() <= 2e-7 * max(1.0, np.abs(p).max()), step
    assert d["weight_decay"] == 0.01 and d["betas"] == (0.9, 0.999)      # what the reference really trains with
