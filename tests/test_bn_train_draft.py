"""CPU: logic check of the DRAFT train-mode BatchNorm kernels (row f1; scouter_b200/csrc/draft/bn_train.cuh, not in the
library) against ``F.batch_norm(training=True)`` + residual + ReLU, by host emulation with a launch-like decomposition."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

DRAFT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scouter_b200", "csrc", "draft")
_f, _d = C.POINTER(C.c_float), C.POINTER(C.c_double)


class Args(C.Structure):            # scouter_draft::BnTrainArgs
    _fields_ = [("M", C.c_longlong), ("C", C.c_int), ("x", _f), ("sums", _d), ("gamma", _f), ("beta", _f),
                ("running_mean", _f), ("running_var", _f), ("scale", _f), ("shift", _f), ("save_mean", _f), ("save_rstd", _f),
                ("eps", C.c_float),
                ("momentum", C.c_float), ("residual", _f), ("y", _f), ("relu", C.c_int)]


class BwdArgs(C.Structure):         # scouter_draft::BnBwdArgs
    _fields_ = [("M", C.c_longlong), ("C", C.c_int), ("x", _f), ("out", _f), ("d_out", _f), ("gamma", _f), ("save_mean", _f),
                ("save_rstd", _f), ("sums", _d), ("d_gamma", _f), ("d_beta", _f), ("coef", _f), ("dx", _f), ("d_residual", _f),
                ("relu", C.c_int)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("bn") / "bn_train_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", os.path.join(DRAFT, "bn_train_host.cpp"),
                    "-o", so], check=True)
    lib = C.CDLL(so)
    lib.bn_train_host.argtypes = [C.POINTER(Args), C.c_int, C.c_int]
    lib.bn_train_host.restype = None
    lib.bn_train_backward_host.argtypes = [C.POINTER(BwdArgs), C.c_int, C.c_int]
    lib.bn_train_backward_host.restype = None
    return lib


@pytest.mark.parametrize("b,c,h,w,relu,res,ctas,threads", [(4, 32, 7, 5, 1, 0, 3, 64), (2, 2048, 3, 3, 1, 1, 5, 256),
                                                           (4, 64, 1, 1, 0, 0, 7, 256), (1, 8, 2, 1, 1, 1, 2, 32),
                                                           (3, 1024, 2, 2, 0, 1, 4, 128)])
def test_bn_train_draft_forward_and_backward_match_torch(emu, b, c, h, w, relu, res, ctas, threads):
    g = torch.Generator().manual_seed(b * 1000 + c)
    x = torch.randn(b, c, h, w, generator=g) * 2 + 5            # |mean| > std: the case fp32 E[x^2]-mean^2 gets wrong
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    rm, rv = torch.randn(c, generator=g), torch.rand(c, generator=g) + 0.5
    r = torch.randn(b, c, h, w, generator=g) if res else None
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ref = F.batch_norm(x, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5)
    if res:
        ref = ref + r
    if relu:
        ref = torch.relu(ref)

    nhwc = lambda t: np.ascontiguousarray(t.permute(0, 2, 3, 1).reshape(-1, c).numpy())
    xs, ys = nhwc(x), np.full((b * h * w, c), np.nan, np.float32)
    rs = nhwc(r) if res else None
    sums = np.zeros((c, 2), np.float64)
    ga, be, rmn, rvn = gamma.numpy().copy(), beta.numpy().copy(), rm.numpy().copy(), rv.numpy().copy()
    scale, shift = np.zeros(c, np.float32), np.zeros(c, np.float32)
    smean, srstd = np.zeros(c, np.float32), np.zeros(c, np.float32)
    p = lambda a_: a_.ctypes.data_as(_f)
    a = Args(M=b * h * w, C=c, x=p(xs), sums=sums.ctypes.data_as(_d), gamma=p(ga), beta=p(be), running_mean=p(rmn),
             running_var=p(rvn), scale=p(scale), shift=p(shift), save_mean=p(smean), save_rstd=p(srstd), eps=1e-5, momentum=0.1,
             residual=p(rs) if res else None, y=p(ys), relu=relu)
    emu.bn_train_host(C.byref(a), ctas, threads)
    got = torch.from_numpy(ys).reshape(b, h, w, c).permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    assert np.allclose(rmn, rm_ref.numpy(), rtol=1e-6, atol=1e-6)
    assert np.allclose(rvn, rv_ref.numpy(), rtol=1e-5, atol=1e-6)
    assert np.allclose(sums[:, 0], xs.astype(np.float64).sum(0), rtol=1e-12)      # every row counted exactly once

    # ---- backward of out = [relu](bn(x) [+ residual]) against autograd -------------------------------------------------
    xa, ga_, ba_ = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ra = r.clone().requires_grad_(True) if res else None
    o = F.batch_norm(xa, None, None, ga_, ba_, True, 0.1, 1e-5)
    if res:
        o = o + ra
    if relu:
        o = torch.relu(o)
    d_out = torch.randn(o.shape, generator=g)
    grads = torch.autograd.grad(o, [xa, ga_, ba_] + ([ra] if res else []), d_out)
    dos, dxs = nhwc(d_out), np.full((b * h * w, c), np.nan, np.float32)
    drs = np.full((b * h * w, c), np.nan, np.float32) if res else None
    sums2, dg, db, coef = np.zeros((c, 2), np.float64), np.zeros(c, np.float32), np.zeros(c, np.float32), np.zeros((c, 3), np.float32)
    bw = BwdArgs(M=b * h * w, C=c, x=p(xs), out=p(ys), d_out=p(dos), gamma=p(ga), save_mean=p(smean), save_rstd=p(srstd),
                 sums=sums2.ctypes.data_as(_d), d_gamma=p(dg), d_beta=p(db), coef=p(coef), dx=p(dxs),
                 d_residual=p(drs) if res else None, relu=relu)
    emu.bn_train_backward_host(C.byref(bw), ctas, threads)
    tol = lambda ref_: 5e-5 * max(1.0, float(ref_.abs().max()))
    got_dx = torch.from_numpy(dxs).reshape(b, h, w, c).permute(0, 3, 1, 2)
    assert float((got_dx - grads[0]).abs().max()) <= tol(grads[0])
    assert float((torch.from_numpy(dg) - grads[1]).abs().max()) <= tol(grads[1])
    assert float((torch.from_numpy(db) - grads[2]).abs().max()) <= tol(grads[2])
    if res:
        got_dr = torch.from_numpy(drs).reshape(b, h, w, c).permute(0, 3, 1, 2)
        assert float((got_dr - grads[3]).abs().max()) <= tol(grads[3])
