"""CPU: logic check of the DRAFT conv weight-gradient kernel (row f1; scouter_b200/csrc/draft/conv_wgrad.cuh, not in the
library) against autograd of ``F.conv2d`` for the conv shapes of the hot path, by host emulation."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

DRAFT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scouter_b200", "csrc", "draft")
_f = C.POINTER(C.c_float)


class Args(C.Structure):            # scouter_draft::WgradArgs
    _fields_ = [(k, C.c_int) for k in ("B", "H", "W", "Cin", "Ho", "Wo", "Cout", "k", "stride", "pad", "groups")] + \
               [("x", _f), ("dy", _f), ("dw", _f), ("db", _f)]


class DArgs(C.Structure):           # scouter_draft::DgradArgs
    _fields_ = [(k, C.c_int) for k in ("B", "H", "W", "Cin", "Ho", "Wo", "Cout", "k", "stride", "pad", "groups")] + \
               [("dy", _f), ("w", _f), ("dx", _f)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("wg") / "conv_wgrad_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", os.path.join(DRAFT, "conv_wgrad_host.cpp"),
                    "-o", so], check=True)
    lib = C.CDLL(so)
    lib.conv_wgrad_host.argtypes = [C.POINTER(Args), C.c_int]
    lib.conv_wgrad_host.restype = None
    lib.conv_dgrad_host.argtypes = [C.POINTER(DArgs)]
    lib.conv_dgrad_host.restype = None
    return lib


# (B, Cin, Cout, H, W, k, stride, pad, groups, bias, splits): 1x1, grouped 3x3 of split attention, deep-stem s2, the
# MNIST stem (1 input channel), resnet18's strided 3x3 and 1x1 shortcut, fc1/fc2 on 1x1 maps
CASES = [(2, 16, 24, 6, 5, 1, 1, 0, 1, 0, 3), (2, 8, 16, 7, 7, 3, 1, 1, 2, 0, 4), (2, 3, 8, 9, 8, 3, 2, 1, 1, 0, 2),
         (3, 1, 4, 10, 10, 3, 2, 1, 1, 0, 5), (2, 8, 12, 8, 8, 3, 2, 1, 1, 0, 1), (2, 8, 12, 7, 7, 1, 2, 0, 1, 0, 2),
         (4, 16, 32, 1, 1, 1, 1, 0, 1, 1, 3)]


@pytest.mark.parametrize("case", CASES)
def test_conv_wgrad_draft_matches_autograd(emu, case):
    b, cin, cout, h, w, k, stride, pad, groups, bias, splits = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin // groups, k, k, generator=g, requires_grad=True)
    bs = torch.randn(cout, generator=g, requires_grad=True) if bias else None
    y = F.conv2d(x, wt, bs, stride, pad, 1, groups)
    dy = torch.randn(y.shape, generator=g)
    grads = torch.autograd.grad(y, [wt] + ([bs] if bias else []), dy)
    nhwc = lambda t: np.ascontiguousarray(t.detach().permute(0, 2, 3, 1).numpy().astype(np.float32))
    xs, dys = nhwc(x), nhwc(dy)
    dw = np.zeros((cout, k, k, cin // groups), np.float32)
    db = np.zeros(cout, np.float32)
    p = lambda a_: a_.ctypes.data_as(_f)
    a = Args(B=b, H=h, W=w, Cin=cin, Ho=y.shape[2], Wo=y.shape[3], Cout=cout, k=k, stride=stride, pad=pad, groups=groups,
             x=p(xs), dy=p(dys), dw=p(dw), db=p(db) if bias else None)
    emu.conv_wgrad_host(C.byref(a), splits)
    ref = grads[0].permute(0, 2, 3, 1)                                   # OIHW -> OHWI
    assert float((torch.from_numpy(dw) - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
    if bias:
        assert float((torch.from_numpy(db) - grads[1]).abs().max()) <= 1e-5 * max(1.0, float(grads[1].abs().max()))


@pytest.mark.parametrize("case", CASES)
def test_conv_dgrad_draft_matches_autograd(emu, case):
    b, cin, cout, h, w, k, stride, pad, groups, _, _ = case
    g = torch.Generator().manual_seed(sum(case) + 1)
    x = torch.randn(b, cin, h, w, generator=g, requires_grad=True)
    wt = torch.randn(cout, cin // groups, k, k, generator=g)
    y = F.conv2d(x, wt, None, stride, pad, 1, groups)
    dy = torch.randn(y.shape, generator=g)
    (dx,) = torch.autograd.grad(y, x, dy)
    dys = np.ascontiguousarray(dy.permute(0, 2, 3, 1).numpy())
    ws = np.ascontiguousarray(wt.permute(0, 2, 3, 1).numpy())
    dxs = np.full((b, h, w, cin), np.nan, np.float32)
    p = lambda a_: a_.ctypes.data_as(_f)
    a = DArgs(B=b, H=h, W=w, Cin=cin, Ho=y.shape[2], Wo=y.shape[3], Cout=cout, k=k, stride=stride, pad=pad, groups=groups,
              dy=p(dys), w=p(ws), dx=p(dxs))
    emu.conv_dgrad_host(C.byref(a))
    got = torch.from_numpy(dxs).permute(0, 3, 1, 2)
    assert float((got - dx).abs().max()) <= 1e-5 * max(1.0, float(dx.abs().max()))
