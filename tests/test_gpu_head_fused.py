"""GPU (-m gpu): the one-kernel head (head_fused.cu) through scouter_head_forward, geometry by geometry, against the
CPU oracle (oracle/head.py:head_forward = sloter/slot_model.py:108-125) on the same seeded inputs.

Covers what the BASELINE configs do not: odd batches (padding CTA of the 2-CTA cluster), R = 128 exactly, one image per
unit, the unit-size fallback for many slots, 1..5 to_k layers, non-square maps.  Tolerance: the reference's own
fp32-vs-fp64 floor on that input (the sum-normalisation amplifies rounding), never below 2e-4 of max|logit|.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import scouter_b200 as sb
from oracle import head as oh
from scouter_b200 import _lib as L
from scouter_b200.plan import split_weights_bf16
from scouter_b200.synth import fill_state_dict

pytestmark = pytest.mark.gpu

# (B, h, w, ch, classes, slots_per_class, to_k_layers, loss_status)
CASES = [
    (5, 7, 7, 2048, 10, 1, 3, -1),    # cfg-3 geometry, odd batch: last unit half empty + padding CTA
    (3, 8, 8, 512, 16, 2, 1, 1),      # n = 64: two images fill the 128-row tile exactly, S = 32
    (2, 10, 10, 256, 7, 1, 5, 1),     # n = 100: one image per unit, five to_k layers (weight ring wraps)
    (4, 7, 7, 1024, 30, 1, 3, 1),     # S = 30: the two-image unit does not fit shared memory -> one image per unit
    (1, 9, 9, 2048, 10, 1, 3, 1),     # batch 1 at the 260^2 geometry
    (7, 4, 5, 128, 3, 2, 2, -1),      # non-square 4x5 map: six images per unit, K = 4 k-blocks
]


def scaled(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_%dx%d_ch%d_C%dx%d_L%d" % c[:7])
def test_fused_head_vs_oracle(case):
    Bn, h, w, ch, ncls, spc, layers, ls = case
    dev = torch.device("cuda", 0)
    n = h * w
    r = np.random.RandomState(1000 + ch + n)
    feat = torch.from_numpy(np.maximum(r.standard_normal((Bn, h, w, ch)), 0).astype(np.float32))       # NHWC
    cw = torch.from_numpy((r.standard_normal((64, ch)) / np.sqrt(ch)).astype(np.float32))
    cb = torch.from_numpy(0.1 * r.standard_normal(64).astype(np.float32))
    m = sb.SlotAttention(ncls, spc, 64, loss_status=ls, power=2, to_k_layer=layers)
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=7))
    sd = {"slot." + k: v.clone() for k, v in m.state_dict().items()}
    sd["conv1x1.weight"] = cw.reshape(64, ch, 1, 1)
    sd["conv1x1.bias"] = cb
    feat_nchw = feat.permute(0, 3, 1, 2).contiguous()
    kw = dict(num_classes=ncls, slots_per_class=spc, loss_status=ls, power=2, return_attn=True)
    ref32 = oh.head_forward(sd, feat_nchw, dtype=torch.float32, **kw)
    ref64 = oh.head_forward(sd, feat_nchw, dtype=torch.float64, **kw)

    m = m.to(dev).eval()
    desc, packed = m.desc_and_pack(dev)
    S = ncls * spc
    fd, wd, bd = feat.to(dev), cw.to(dev), cb.to(dev)
    wsplit = split_weights_bf16(wd)
    pe = sb.build_position_encoding("sine", 64).table(h, w, dev)
    logits = torch.full((Bn, ncls), float("nan"), device=dev)
    attn = torch.full((Bn, S, n), float("nan"), device=dev)
    asum = torch.full((Bn,), float("nan"), device=dev)
    xo = torch.full((Bn, n, 64), float("nan"), device=dev)
    io = L.HeadIO()
    io.batch, io.h, io.w, io.channel, io.layout, io.math = Bn, h, w, ch, L.LAYOUT_NHWC, L.MATH_TC
    io.feat, io.conv_w, io.conv_b, io.pe = fd.data_ptr(), wd.data_ptr(), bd.data_ptr(), pe.data_ptr()
    io.logits, io.attn, io.attn_sum, io.x_out = logits.data_ptr(), attn.data_ptr(), asum.data_ptr(), xo.data_ptr()
    io.conv_w_split = wsplit.data_ptr()
    lib = L.lib()
    assert lib.scouter_head_launch_count(C.byref(desc), C.byref(io)) == 1, "this geometry must take the fused kernel"
    nbytes = lib.scouter_head_workspace_bytes(C.byref(desc), C.byref(io))
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 1024
    for _ in range(2):      # twice: the second run must reproduce the first bit for bit
        L.check(lib.scouter_head_forward(C.byref(desc), packed.data_ptr(), C.byref(io), ws.data_ptr() + off, nbytes, 0))
        torch.cuda.synchronize()
        if _ == 0:
            first = (logits.clone(), attn.clone())
    assert torch.equal(first[0], logits) and torch.equal(first[1], attn)
    assert not torch.isnan(logits).any() and not torch.isnan(attn).any() and not torch.isnan(xo).any()

    x_ref = torch.relu(feat.double().reshape(Bn * n, ch) @ cw.double().t() + cb.double()).reshape(Bn, n, 64)
    assert scaled(xo, x_ref) < 2e-5                                           # projection: fp32-class on the tensor cores
    floor = scaled(ref32["logits"], ref64["logits"])
    err = scaled(logits, ref64["logits"])
    afloor = float((ref32["attn"].double() - ref64["attn"]).abs().max())
    aerr = float((attn.cpu().double() - ref64["attn"]).abs().max())
    print(f"fused head {case}: logits err {err:.2e} (reference fp32 floor {floor:.2e}), attn err {aerr:.2e} (floor {afloor:.2e})")
    assert err < max(2e-4, 20 * floor)
    assert aerr < max(2e-4, 20 * afloor)
    assert torch.allclose(asum.cpu().double(), ref64["attn"].sum((1, 2)), rtol=1e-4, atol=1e-3)
