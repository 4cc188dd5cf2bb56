"""CPU: logic check of the DRAFT CUDA head backward (row f1; scouter_b200/csrc/draft/, not in the library).

The kernel body is written so that it compiles as host code (one "thread" per image, block barriers become no-ops,
atomics plain adds); this test builds that emulation with g++ and compares every output with the explicit-formula
oracle (oracle/head_backward.py, itself equal to autograd).  It says the transcription is right; it says nothing about
races or performance on a GPU -- the kernel has not run on one yet."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

import scouter_b200 as sb
from oracle import head as oh
from oracle.head_backward import head_backward
from scouter_b200.synth import fill_state_dict

HERE = os.path.dirname(os.path.abspath(__file__))
DRAFT = os.path.join(os.path.dirname(HERE), "scouter_b200", "csrc", "draft")
MAX_L = 8
_f = C.POINTER(C.c_float)


class Args(C.Structure):            # scouter_draft::HeadBwdArgs
    _fields_ = [(k, C.c_int) for k in ("B", "n", "ch", "S", "C", "spc", "L", "iters", "loss_status")] + \
               [("feat", _f), ("conv_w", _f), ("conv_b", _f), ("pe", _f), ("to_k_w", _f * MAX_L), ("to_k_b", _f * MAX_L),
                ("w_ih", _f), ("w_hh", _f), ("b_ih", _f), ("b_hh", _f), ("slots0", _f), ("g_logits", _f), ("attn_coef", _f),
                ("d_feat", _f), ("d_pre", _f), ("g_conv_w", _f), ("g_conv_b", _f), ("g_to_k_w", _f * MAX_L), ("g_to_k_b", _f * MAX_L),
                ("g_w_ih", _f), ("g_w_hh", _f), ("g_b_ih", _f), ("g_b_hh", _f), ("g_slots0", _f),
                ("scratch", _f), ("scratch_per_image", C.c_size_t)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("hb") / "head_backward_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                    os.path.join(DRAFT, "head_backward_host.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    lib.head_backward_scratch_floats.restype = C.c_size_t
    lib.head_backward_scratch_floats.argtypes = [C.c_int] * 4
    lib.head_backward_host.argtypes = [C.POINTER(Args), C.c_int]
    lib.head_backward_host.restype = None
    return lib


def ptr(a):
    return a.ctypes.data_as(_f)


@pytest.mark.parametrize("case", [(10, 1, 3, 1, 2, 3, 64, 7, 7), (5, 2, 1, -1, 1, 2, 32, 3, 4), (30, 1, 3, 1, 2, 2, 48, 9, 9),
                                  (4, 3, 2, -1, 3, 1, 16, 2, 2)])
def test_draft_kernel_body_matches_oracle(emu, case):
    Cn, spc, L, ls, pw, B, ch, h, w = case
    n, S = h * w, Cn * spc
    m = torch.nn.Module()                      # the head's parameters under the reference's key names
    m.conv1x1 = torch.nn.Conv2d(ch, 64, 1)
    m.slot = sb.SlotAttention(Cn, spc, 64, loss_status=ls, power=pw, to_k_layer=L)
    sd = fill_state_dict(m.state_dict(), seed=2)
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(B, ch, h, w, generator=g).abs()
    g_logits = torch.randn(B, Cn, generator=g)
    g_attn = 0.7
    ref = head_backward(sd, feat, g_logits, g_attn, num_classes=Cn, slots_per_class=spc, loss_status=ls, power=pw)
    fwd = oh.head_forward(sd, feat, num_classes=Cn, slots_per_class=spc, loss_status=ls, power=pw, dtype=torch.float64,
                          return_attn=True)
    mean = float(fwd["attn"].sum()) / (B * S * n)
    coef = np.array([g_attn * pw * mean ** (pw - 1) / (B * S * n)], np.float32)

    f32 = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np.float32))
    keep = dict(feat=f32(feat.reshape(B, ch, n).permute(0, 2, 1)), conv_w=f32(sd["conv1x1.weight"].reshape(64, ch)),
                conv_b=f32(sd["conv1x1.bias"]), pe=f32(oh.sine_pe(64, h, w).reshape(64, n).t()),
                w_ih=f32(sd["slot.gru.weight_ih_l0"]), w_hh=f32(sd["slot.gru.weight_hh_l0"]),
                b_ih=f32(sd["slot.gru.bias_ih_l0"]), b_hh=f32(sd["slot.gru.bias_hh_l0"]),
                slots0=f32(sd["slot.initial_slots"][0]), g_logits=f32(g_logits), attn_coef=coef)
    out = dict(d_feat=np.zeros((B, n, ch), np.float32), g_conv_w=np.zeros((64, ch), np.float32), g_conv_b=np.zeros(64, np.float32),
               g_w_ih=np.zeros((192, 64), np.float32), g_w_hh=np.zeros((192, 64), np.float32), g_b_ih=np.zeros(192, np.float32),
               g_b_hh=np.zeros(192, np.float32), g_slots0=np.zeros((S, 64), np.float32))
    a = Args(B=B, n=n, ch=ch, S=S, C=Cn, spc=spc, L=L, iters=3, loss_status=ls)
    for k, v in {**keep, **out}.items():
        setattr(a, k, ptr(v))
    kw, kb, gkw, gkb = [], [], [], []
    for l in range(L):
        kw.append(f32(sd[f"slot.to_k.{2 * l}.weight"])); kb.append(f32(sd[f"slot.to_k.{2 * l}.bias"]))
        gkw.append(np.zeros((64, 64), np.float32)); gkb.append(np.zeros(64, np.float32))
        a.to_k_w[l], a.to_k_b[l], a.g_to_k_w[l], a.g_to_k_b[l] = ptr(kw[l]), ptr(kb[l]), ptr(gkw[l]), ptr(gkb[l])
    per = emu.head_backward_scratch_floats(n, S, L, 3)
    scratch = np.full(B * per, np.nan, np.float32)          # reads of anything never written would poison the result
    a.scratch, a.scratch_per_image = ptr(scratch), per
    emu.head_backward_host(C.byref(a), 1)

    got = {"feat": torch.from_numpy(out["d_feat"]).permute(0, 2, 1).reshape(B, ch, h, w),
           "conv1x1.weight": torch.from_numpy(out["g_conv_w"]).reshape(64, ch, 1, 1), "conv1x1.bias": torch.from_numpy(out["g_conv_b"]),
           "slot.gru.weight_ih_l0": torch.from_numpy(out["g_w_ih"]), "slot.gru.weight_hh_l0": torch.from_numpy(out["g_w_hh"]),
           "slot.gru.bias_ih_l0": torch.from_numpy(out["g_b_ih"]), "slot.gru.bias_hh_l0": torch.from_numpy(out["g_b_hh"]),
           "slot.initial_slots": torch.from_numpy(out["g_slots0"])[None]}
    for l in range(L):
        got[f"slot.to_k.{2 * l}.weight"], got[f"slot.to_k.{2 * l}.bias"] = torch.from_numpy(gkw[l]), torch.from_numpy(gkb[l])
    assert set(got) == set(ref)
    got = {k: v.clone() for k, v in got.items()}             # the arrays are reused by the second run
    # second mode: the kernel stops at d_pre (B, n, 64); the two ch-sized products are then plain GEMMs
    d_pre = np.full((B, n, 64), np.nan, np.float32)
    for v in list(out.values()) + gkw + gkb:
        v[...] = 0
    scratch[...] = np.nan
    a.d_feat, a.d_pre = None, ptr(d_pre)
    emu.head_backward_host(C.byref(a), 1)
    dp = torch.from_numpy(d_pre).double()
    tokens = torch.from_numpy(keep["feat"]).double()
    wc = torch.from_numpy(keep["conv_w"]).double()
    via_gemm = {"feat": (dp @ wc).permute(0, 2, 1).reshape(B, ch, h, w),
                "conv1x1.weight": (dp.reshape(-1, 64).t() @ tokens.reshape(-1, ch)).reshape(64, ch, 1, 1),
                "conv1x1.bias": torch.from_numpy(out["g_conv_b"]).double()}
    assert not out["g_conv_w"].any() and not out["d_feat"].any()            # untouched in this mode
    for k, r in ref.items():
        if k in via_gemm:
            e2 = float((via_gemm[k] - r).abs().max() / r.abs().max().clamp_min(1e-30))
            assert e2 < 5e-4, (k, e2)
        err = float((got[k].double() - r).abs().max() / r.abs().max().clamp_min(1e-30))
        assert err < 5e-4, (k, err)       # fp32 kernel vs fp64 formulas (measured: <= 3e-5 on these cases)
