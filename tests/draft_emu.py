"""Test helper: numpy wrappers around the row-f1 DRAFT kernels (scouter_b200/csrc/draft/).  Everything is NHWC float32 numpy
in, numpy out.  Two backends run the SAME kernel bodies from the SAME ctypes argument blocks:

* ``host`` (default) -- the *_host.cpp emulations, built once with g++ into one shared object (CPU tests);
* ``gpu``            -- ``draft_api.cu`` + the draft ``.cu`` files built with nvcc into their own ``libscouter_draft.so``
                        (never into libscouter_b200.so); every array handed to a kernel is uploaded, the kernel is
                        launched on the current device, and the arrays are copied back.  Selected with
                        ``SCOUTER_DRAFT_BACKEND=gpu`` or ``draft_emu.BACKEND = "gpu"``; used by ``pytest -m gpu_draft``."""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

from test_bn_train_draft import Args as BnArgs, BwdArgs as BnBwdArgs
from test_conv_wgrad_draft import Args as WgArgs, DArgs as DgArgs
from test_head_backward_draft import Args as HeadArgs
from test_pool_splat_bwd_draft import PoolArgs, SplatArgs

DRAFT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scouter_b200", "csrc", "draft")
_f, _d = C.POINTER(C.c_float), C.POINTER(C.c_double)
_lib = None
_gpu_lib = None
BACKEND = os.environ.get("SCOUTER_DRAFT_BACKEND", "host")
_dev = []           # (numpy array, its device copy) of the call in flight


def lib():
    global _lib
    if _lib is None:
        if not shutil.which("g++"):
            return None
        so = os.path.join(tempfile.mkdtemp(prefix="scouter_draft_"), "draft_host.so")
        srcs = [os.path.join(DRAFT, f) for f in sorted(os.listdir(DRAFT)) if f.endswith("_host.cpp")]
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", *srcs, "-o", so], check=True)
        _lib = C.CDLL(so)
        _lib.head_backward_scratch_floats.restype = C.c_size_t
        _lib.head_backward_scratch_floats.argtypes = [C.c_int] * 4
    return _lib


def gpu_lib():
    global _gpu_lib
    if _gpu_lib is None:
        so = os.path.join(tempfile.mkdtemp(prefix="scouter_draft_gpu_"), "libscouter_draft.so")
        srcs = [os.path.join(DRAFT, f) for f in sorted(os.listdir(DRAFT)) if f.endswith(".cu")]
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                        "-shared", *srcs, "-o", so], check=True, capture_output=True)
        _gpu_lib = C.CDLL(so)
        _gpu_lib.draft_head_backward_scratch_floats.restype = C.c_size_t
        _gpu_lib.draft_head_backward_scratch_floats.argtypes = [C.c_int] * 4
    return _gpu_lib


def _ptr(a, ctype):
    if a is None:
        return None
    if BACKEND == "gpu":
        import torch
        assert a.flags["C_CONTIGUOUS"]
        t = torch.from_numpy(a).cuda()
        _dev.append((a, t))
        return C.cast(t.data_ptr(), ctype)
    return a.ctypes.data_as(ctype)


def p(a):
    return _ptr(a, _f)


def pd(a):
    return _ptr(a, _d)


def _call(host_name, gpu_name, args, host_extra=(), gpu_extra=()):
    """Run one draft kernel (group) on the selected backend; on the GPU, check the launch and copy every array back."""
    if BACKEND == "gpu":
        import torch
        rc = getattr(gpu_lib(), gpu_name)(C.byref(args), *[C.c_int(v) for v in gpu_extra])
        torch.cuda.synchronize()
        assert rc == 0, (gpu_name, rc)
        for a, t in _dev:
            a[...] = t.cpu().numpy()
        _dev.clear()
    else:
        getattr(lib(), host_name)(C.byref(args), *[C.c_int(v) for v in host_extra])


def c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def bn_train_forward(x, gamma, beta, rm, rv, residual=None, relu=True, ctas=3, threads=64):
    """x (..., C) -> (y, save_mean, save_rstd); rm / rv updated in place."""
    x = c32(x)
    cch = x.shape[-1]
    m = x.size // cch
    y = np.empty_like(x)
    sums = np.zeros((cch, 2), np.float64)
    scale, shift, mean, rstd = (np.zeros(cch, np.float32) for _ in range(4))
    res = None if residual is None else c32(residual)
    a = BnArgs(M=m, C=cch, x=p(x), sums=pd(sums), gamma=p(gamma), beta=p(beta), running_mean=p(rm), running_var=p(rv),
               scale=p(scale), shift=p(shift), save_mean=p(mean), save_rstd=p(rstd), eps=1e-5, momentum=0.1, residual=p(res),
               y=p(y), relu=int(relu))
    _call("bn_train_host", "draft_bn_train", a, (ctas, threads))
    return y, mean, rstd


def bn_train_backward(x, out, d_out, gamma, mean, rstd, relu=True, want_residual=False, ctas=3, threads=64):
    """-> (dx, d_gamma, d_beta, d_residual or None)"""
    x, out, d_out = c32(x), c32(out), c32(d_out)
    cch = x.shape[-1]
    m = x.size // cch
    dx = np.empty_like(x)
    dres = np.empty_like(x) if want_residual else None
    sums = np.zeros((cch, 2), np.float64)
    dg, db, coef = np.zeros(cch, np.float32), np.zeros(cch, np.float32), np.zeros((cch, 3), np.float32)
    a = BnBwdArgs(M=m, C=cch, x=p(x), out=p(out), d_out=p(d_out), gamma=p(gamma), save_mean=p(mean), save_rstd=p(rstd),
                  sums=pd(sums), d_gamma=p(dg), d_beta=p(db), coef=p(coef), dx=p(dx), d_residual=p(dres), relu=int(relu))
    _call("bn_train_backward_host", "draft_bn_train_backward", a, (ctas, threads))
    return dx, dg, db, dres


def conv_backward(x, dy, w_ohwi, stride, pad, groups, bias=False, need_dx=True, splits=3):
    """x (B,H,W,Cin), dy (B,Ho,Wo,Cout), w (Cout,k,k,Cin/g) -> (dx or None, dw, db or None)"""
    x, dy, w_ohwi = c32(x), c32(dy), c32(w_ohwi)
    b, h, w, cin = x.shape
    _, ho, wo, cout = dy.shape
    k = w_ohwi.shape[1]
    dw = np.zeros_like(w_ohwi)
    db = np.zeros(cout, np.float32) if bias else None
    a = WgArgs(B=b, H=h, W=w, Cin=cin, Ho=ho, Wo=wo, Cout=cout, k=k, stride=stride, pad=pad, groups=groups, x=p(x), dy=p(dy),
               dw=p(dw), db=p(db))
    _call("conv_wgrad_host", "draft_conv_wgrad", a, (splits,))
    dx = None
    if need_dx:
        dx = np.empty_like(x)
        d = DgArgs(B=b, H=h, W=w, Cin=cin, Ho=ho, Wo=wo, Cout=cout, k=k, stride=stride, pad=pad, groups=groups, dy=p(dy), w=p(w_ohwi),
                   dx=p(dx))
        _call("conv_dgrad_host", "draft_conv_dgrad", d)
    return dx, dw, db


def pool_backward(kind, x, dy):
    """kind 0 max(3,2,1) / 1 avg(2,2,ceil) / 2 avg(3,2,1); x (B,H,W,C) forward input, dy (B,Ho,Wo,C)"""
    x, dy = c32(x), c32(dy)
    b, h, w, cch = x.shape
    dx = np.empty_like(x)
    a = PoolArgs(B=b, H=h, W=w, C=cch, Ho=dy.shape[1], Wo=dy.shape[2], x=p(x), dy=p(dy), dx=p(dx))
    _call("pool_bwd_host", "draft_pool_bwd", a, (kind,), (kind,))
    return dx


def splat_backward_logits(x2, d_out, att):
    """x2 (B,H,W,2C), d_out (B,H,W,C), att (B,2,C) -> d_logit (B,2,C)"""
    x2, d_out, att = c32(x2), c32(d_out), c32(att)
    b, h, w, c2 = x2.shape
    cch = c2 // 2
    d_att, d_logit = np.empty((b, 2, cch), np.float32), np.empty((b, 2, cch), np.float32)
    a = SplatArgs(B=b, HW=h * w, C=cch, x2=p(x2), d_out=p(d_out), att=p(att), d_att=p(d_att), d_logit=p(d_logit), d_gap=None, d_x2=None)
    _call("splat_bwd_host", "draft_splat_bwd", a, (0,), (0,))
    return d_logit


def splat_backward_apply(x2, d_out, att, d_gap):
    x2, d_out, att, d_gap = c32(x2), c32(d_out), c32(att), c32(d_gap)
    b, h, w, c2 = x2.shape
    d_x2 = np.empty_like(x2)
    a = SplatArgs(B=b, HW=h * w, C=c2 // 2, x2=p(x2), d_out=p(d_out), att=p(att), d_att=None, d_logit=None, d_gap=p(d_gap), d_x2=p(d_x2))
    _call("splat_bwd_host", "draft_splat_bwd", a, (1,), (1,))
    return d_x2


def head_backward(feat_tokens, sd, pe, g_logits, attn_coef, num_classes, spc, loss_status, n_layers):
    """feat_tokens (B,n,ch) -> (d_feat (B,n,ch), {state_dict key: gradient})"""
    f = c32(feat_tokens)
    b, n, ch = f.shape
    s = num_classes * spc
    keep = dict(conv_w=c32(sd["conv1x1.weight"].reshape(64, ch)), conv_b=c32(sd["conv1x1.bias"]), pe=c32(pe),
                w_ih=c32(sd["slot.gru.weight_ih_l0"]), w_hh=c32(sd["slot.gru.weight_hh_l0"]), b_ih=c32(sd["slot.gru.bias_ih_l0"]),
                b_hh=c32(sd["slot.gru.bias_hh_l0"]), slots0=c32(sd["slot.initial_slots"][0]), g_logits=c32(g_logits),
                attn_coef=np.array([attn_coef], np.float32))
    out = dict(d_feat=np.zeros((b, n, ch), np.float32), g_conv_w=np.zeros((64, ch), np.float32), g_conv_b=np.zeros(64, np.float32),
               g_w_ih=np.zeros((192, 64), np.float32), g_w_hh=np.zeros((192, 64), np.float32), g_b_ih=np.zeros(192, np.float32),
               g_b_hh=np.zeros(192, np.float32), g_slots0=np.zeros((s, 64), np.float32))
    a = HeadArgs(B=b, n=n, ch=ch, S=s, C=num_classes, spc=spc, L=n_layers, iters=3, loss_status=loss_status, feat=p(f))
    for k, v in {**keep, **out}.items():
        setattr(a, k, p(v))
    a.d_pre = None
    kw, kb, gkw, gkb = [], [], [], []
    for l in range(n_layers):
        kw.append(c32(sd[f"slot.to_k.{2 * l}.weight"])); kb.append(c32(sd[f"slot.to_k.{2 * l}.bias"]))
        gkw.append(np.zeros((64, 64), np.float32)); gkb.append(np.zeros(64, np.float32))
        a.to_k_w[l], a.to_k_b[l], a.g_to_k_w[l], a.g_to_k_b[l] = p(kw[l]), p(kb[l]), p(gkw[l]), p(gkb[l])
    per = (gpu_lib().draft_head_backward_scratch_floats if BACKEND == "gpu" else lib().head_backward_scratch_floats)(n, s, n_layers, 3)
    scratch = np.full(b * per, np.nan, np.float32)
    a.scratch, a.scratch_per_image = p(scratch), per
    _call("head_backward_host", "draft_head_backward", a, (1,))
    grads = {"conv1x1.weight": out["g_conv_w"].reshape(64, ch, 1, 1), "conv1x1.bias": out["g_conv_b"],
             "slot.gru.weight_ih_l0": out["g_w_ih"], "slot.gru.weight_hh_l0": out["g_w_hh"], "slot.gru.bias_ih_l0": out["g_b_ih"],
             "slot.gru.bias_hh_l0": out["g_b_hh"], "slot.initial_slots": out["g_slots0"][None]}
    for l in range(n_layers):
        grads[f"slot.to_k.{2 * l}.weight"], grads[f"slot.to_k.{2 * l}.bias"] = gkw[l], gkb[l]
    return out["d_feat"], grads


def adamw_step(params, grads, exp_avg, exp_avg_sq, lr, step, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    """In-place AdamW step (torch.optim.AdamW semantics) over flat float32 arrays."""
    import math
    from test_adamw_draft import Args as AdamArgs
    b1, b2 = betas
    a = AdamArgs(decay=1 - lr * weight_decay, one_minus_beta1=1 - b1, beta2=b2, one_minus_beta2=1 - b2, eps=eps,
                 step_size=lr / (1 - b1 ** step), bias_correction2_sqrt=math.sqrt(1 - b2 ** step), n=params.size, p=p(params),
                 g=p(grads), m=p(exp_avg), v=p(exp_avg_sq))
    _call("adamw_host", "draft_adamw", a)
