"""CPU: logic check of the DRAFT pooling / split-attention backward (row f1; scouter_b200/csrc/draft/pool_splat_bwd.cuh,
not in the library) against autograd of the same torch ops the reference uses, by host emulation."""
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


class PoolArgs(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("B", "H", "W", "C", "Ho", "Wo")] + [("x", _f), ("dy", _f), ("dx", _f)]


class SplatArgs(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("B", "HW", "C")] + [(k, _f) for k in ("x2", "d_out", "att", "d_att", "d_logit", "d_gap", "d_x2")]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("ps") / "pool_splat_bwd_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                    os.path.join(DRAFT, "pool_splat_bwd_host.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    lib.pool_bwd_host.argtypes = [C.POINTER(PoolArgs), C.c_int]
    lib.splat_bwd_host.argtypes = [C.POINTER(SplatArgs), C.c_int]
    lib.pool_bwd_host.restype = lib.splat_bwd_host.restype = None
    return lib


def nhwc(t):
    return np.ascontiguousarray(t.detach().permute(0, 2, 3, 1).numpy().astype(np.float32))


def p(a):
    return a.ctypes.data_as(_f)


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("b,c,h,w", [(2, 4, 8, 8), (1, 8, 7, 9), (2, 4, 5, 5), (1, 4, 1, 2)])
def test_pool_backward_draft(emu, kind, b, c, h, w):
    g = torch.Generator().manual_seed(kind * 100 + h * 10 + w)
    x = torch.relu(torch.randn(b, c, h, w, generator=g)).requires_grad_(True)      # post-ReLU input: many exact ties at 0
    y = [lambda t: F.max_pool2d(t, 3, 2, 1),
         lambda t: F.avg_pool2d(t, 2, 2, ceil_mode=True, count_include_pad=False),
         lambda t: F.avg_pool2d(t, 3, 2, 1)][kind](x)
    dy = torch.randn(y.shape, generator=g)
    (dx,) = torch.autograd.grad(y, x, dy)
    xs, dys, dxs = nhwc(x), nhwc(dy), np.full((b, h, w, c), np.nan, np.float32)
    a = PoolArgs(B=b, H=h, W=w, C=c, Ho=y.shape[2], Wo=y.shape[3], x=p(xs), dy=p(dys), dx=p(dxs))
    emu.pool_bwd_host(C.byref(a), kind)
    got = torch.from_numpy(dxs).permute(0, 3, 1, 2)
    assert float((got - dx).abs().max()) <= 1e-6 * max(1.0, float(dx.abs().max()))


@pytest.mark.parametrize("b,c,h,w,mid", [(2, 8, 5, 4, 32), (3, 16, 1, 1, 32), (1, 4, 7, 7, 32)])
def test_split_attention_backward_draft(emu, b, c, h, w, mid):
    """reduce + softmax stages against autograd's gradient at fc2's output, then -- with d_gap taken from autograd of
    the fc1/bn1/fc2 chain -- the apply stage against the full gradient at the split-attention input."""
    g = torch.Generator().manual_seed(b * 10 + c)
    x2 = torch.relu(torch.randn(b, 2 * c, h, w, generator=g)).requires_grad_(True)
    w1, b1 = torch.randn(mid, c, 1, 1, generator=g) * 0.3, torch.randn(mid, generator=g) * 0.1
    w2, b2 = torch.randn(2 * c, mid, 1, 1, generator=g) * 0.3, torch.randn(2 * c, generator=g) * 0.1
    xr = x2.reshape(b, 2, c, h, w)
    gap = xr.sum(1).mean((2, 3), keepdim=True)
    gap.retain_grad()
    logits = F.conv2d(torch.relu(F.conv2d(gap, w1, b1)), w2, b2)                  # (B, 2C, 1, 1) radix-major (bn1 omitted: any map)
    logits.retain_grad()
    att = torch.softmax(logits.reshape(b, 2, c), dim=1)
    out = (xr * att[:, :, :, None, None]).sum(1)
    d_out = torch.randn(out.shape, generator=g)
    out.backward(d_out)

    f32 = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np.float32))
    x2s, dos = nhwc(x2).reshape(b, h * w, 2 * c), nhwc(d_out).reshape(b, h * w, c)
    atts = f32(att)
    d_att, d_logit = np.full((b, 2, c), np.nan, np.float32), np.full((b, 2, c), np.nan, np.float32)
    d_gap = f32(gap.grad.reshape(b, c))
    d_x2 = np.full((b, h * w, 2 * c), np.nan, np.float32)
    a = SplatArgs(B=b, HW=h * w, C=c, x2=p(x2s), d_out=p(dos), att=p(atts), d_att=p(d_att), d_logit=p(d_logit), d_gap=p(d_gap),
                  d_x2=p(d_x2))
    emu.splat_bwd_host(C.byref(a), 0)
    ref_logit = logits.grad.reshape(b, 2, c)
    assert float((torch.from_numpy(d_logit) - ref_logit).abs().max()) <= 2e-5 * max(1.0, float(ref_logit.abs().max()))
    emu.splat_bwd_host(C.byref(a), 1)
    got = torch.from_numpy(d_x2).reshape(b, h, w, 2 * c).permute(0, 3, 1, 2)
    assert float((got - x2.grad).abs().max()) <= 2e-5 * max(1.0, float(x2.grad.abs().max()))
