"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the backward of the xSlot head written out by hand.

Row f1 groundwork.  ``oracle/train.py`` gets the reference's gradients from autograd -- that is the oracle.  This file
is the same backward as explicit formulas (no autograd), in the order a fused CUDA kernel would evaluate them
(re-compute the forward per image, walk the three attention iterations backwards, two GRU steps of BPTT), checked
against autograd in ``tests/test_oracle_golden.py``.  It documents the math the CUDA backward has to implement:

forward (slot_model.py:108-116, slot_attention.py:44-96), per image, S slots, n tokens, d = 64::

    x   = relu(conv1x1(feat))                 (n,d)        k = to_k(x + pe)          (n,d)
    for t in 0,1,2:
        dots = s_t k^T * d^-1/2               (S,n)
        row_i = sum_j dots_ij ;  tot = sum_ij dots_ij
        attn = sigmoid(dots / row * tot)                    # no eps (slot_attention.py:56-57)
        upd_t = attn x / d                    (S,d)         # divides by d, not n (:58-59)
        s_{t+1} = GRU(upd_t, s_t)                           # s_3 is dead (:60-66 on the last iteration)
    logits_c = loss_status * sum_{slots of c, e} upd_2      attn_loss = (sum attn_2 / (B S n))^power

Everything is float tensors in, float tensors out; ``dtype`` picks fp64 for the comparison with autograd.
"""
from __future__ import annotations

import torch

from .head import sine_pe


def _gru_forward(u, s, w_ih, w_hh, b_ih, b_hh):
    d = s.shape[-1]
    gi = u @ w_ih.t() + b_ih
    gh = s @ w_hh.t() + b_hh
    r = torch.sigmoid(gi[..., :d] + gh[..., :d])
    z = torch.sigmoid(gi[..., d:2 * d] + gh[..., d:2 * d])
    nn = torch.tanh(gi[..., 2 * d:] + r * gh[..., 2 * d:])
    return (1 - z) * nn + z * s, (r, z, nn, gh[..., 2 * d:])


def _gru_backward(ds_new, u, s, cache, w_ih, w_hh):
    """d(loss)/d(u, s, W_ih, W_hh, b_ih, b_hh) of one GRU step given d(loss)/d(s_new)."""
    r, z, nn, gh_n = cache
    dnn = ds_new * (1 - z)
    dz = ds_new * (s - nn)
    ds = ds_new * z
    da_n = dnn * (1 - nn * nn)
    dr = da_n * gh_n
    da_r = dr * r * (1 - r)
    da_z = dz * z * (1 - z)
    dgi = torch.cat([da_r, da_z, da_n], -1)                 # gate order [r | z | n]
    dgh = torch.cat([da_r, da_z, da_n * r], -1)
    du = dgi @ w_ih
    ds = ds + dgh @ w_hh
    flat = lambda t: t.reshape(-1, t.shape[-1])
    return du, ds, flat(dgi).t() @ flat(u), flat(dgh).t() @ flat(s), flat(dgi).sum(0), flat(dgh).sum(0)


def head_backward(sd: dict, feat: torch.Tensor, g_logits: torch.Tensor, g_attn_loss, *, num_classes: int,
                  slots_per_class: int, loss_status: int = 1, power: int = 1, iters: int = 3, dtype=torch.float64):
    """Gradients of ``sum(g_logits * logits) + g_attn_loss * attn_loss`` w.r.t. the backbone features and every head
    parameter.  ``feat``: (B, ch, h, w).  Returns {"feat": ..., "conv1x1.weight": ..., "slot.gru.weight_ih_l0": ...}."""
    g = lambda k: sd[k].detach().to(dtype)
    feat = feat.detach().to(dtype)
    g_logits = g_logits.to(dtype)
    b, ch, h, w_ = feat.shape
    n = h * w_
    wc, bc = g("conv1x1.weight").reshape(-1, ch), g("conv1x1.bias")
    d = wc.shape[0]
    S = num_classes * slots_per_class
    scale = d ** -0.5

    # ---- forward, keeping what the backward needs -------------------------------------------------------------------
    f = feat.reshape(b, ch, n).permute(0, 2, 1)                              # (B,n,ch) token-major
    x = torch.relu(f @ wc.t() + bc)                                          # (B,n,d)
    pe = sine_pe(d, h, w_, dtype).reshape(d, n).t()
    lin = []
    li = 0
    while f"slot.to_k.{li}.weight" in sd:
        lin.append((g(f"slot.to_k.{li}.weight"), g(f"slot.to_k.{li}.bias"), li))
        li += 2
    acts = [x + pe]                                                          # inputs of each Linear
    for i, (wl, bl, _) in enumerate(lin):
        a = acts[-1] @ wl.t() + bl
        acts.append(torch.relu(a) if i + 1 < len(lin) else a)
    k = acts[-1]
    w_ih, w_hh = g("slot.gru.weight_ih_l0"), g("slot.gru.weight_hh_l0")
    b_ih, b_hh = g("slot.gru.bias_ih_l0"), g("slot.gru.bias_hh_l0")
    s = [g("slot.initial_slots").expand(b, -1, -1)]
    dots_, attn_, upd_, gru_ = [], [], [], []
    for t in range(iters):
        dots = s[t] @ k.transpose(1, 2) * scale
        row = dots.sum(2, keepdim=True)
        tot = dots.sum((1, 2), keepdim=True)
        attn = torch.sigmoid(dots / row * tot)
        upd = attn @ x / d
        dots_.append(dots); attn_.append(attn); upd_.append(upd)
        if t + 1 < iters:                                                    # the last GRU step is dead
            s_new, cache = _gru_forward(upd, s[t], w_ih, w_hh, b_ih, b_hh)
            s.append(s_new); gru_.append(cache)

    # ---- backward ---------------------------------------------------------------------------------------------------
    m = attn_[-1].sum() / (b * S * n)
    d_attn_last = g_attn_loss * power * m ** (power - 1) / (b * S * n)      # relu(attn) == attn (:92-94)
    d_upd = (loss_status * g_logits).repeat_interleave(slots_per_class, dim=1)[:, :, None].expand(b, S, d)
    d_s_next = torch.zeros(b, S, d, dtype=dtype)
    d_x = torch.zeros_like(x)
    d_k = torch.zeros_like(k)
    gw_ih, gw_hh = torch.zeros_like(w_ih), torch.zeros_like(w_hh)
    gb_ih, gb_hh = torch.zeros_like(b_ih), torch.zeros_like(b_hh)
    for t in reversed(range(iters)):
        d_s = torch.zeros(b, S, d, dtype=dtype)
        if t + 1 < iters:                                                    # through s_{t+1} = GRU(upd_t, s_t)
            d_upd, d_s, a1, a2, a3, a4 = _gru_backward(d_s_next, upd_[t], s[t], gru_[t], w_ih, w_hh)
            gw_ih += a1; gw_hh += a2; gb_ih += a3; gb_hh += a4
        attn, dots = attn_[t], dots_[t]
        d_attn = d_upd @ x.transpose(1, 2) / d                               # upd = attn x / d
        if t + 1 == iters:
            d_attn = d_attn + d_attn_last
        d_x += attn.transpose(1, 2) @ d_upd / d
        du = d_attn * attn * (1 - attn)                                      # sigmoid
        row = dots.sum(2, keepdim=True)
        tot = dots.sum((1, 2), keepdim=True)
        d_row = -(du * dots).sum(2, keepdim=True) * tot / (row * row)        # u = dots * tot / row
        d_tot = (du * dots / row).sum((1, 2), keepdim=True)
        d_dots = du * tot / row + d_row + d_tot
        d_s = d_s + d_dots @ k * scale                                       # dots = s k^T * scale
        d_k += d_dots.transpose(1, 2) @ s[t] * scale
        d_s_next = d_s
    out = {"slot.initial_slots": d_s_next.sum(0, keepdim=True),
           "slot.gru.weight_ih_l0": gw_ih, "slot.gru.weight_hh_l0": gw_hh,
           "slot.gru.bias_ih_l0": gb_ih, "slot.gru.bias_hh_l0": gb_hh}
    d_a = d_k                                                                # to_k MLP, last Linear has no ReLU
    for i in reversed(range(len(lin))):
        wl, _, idx = lin[i]
        if i + 1 < len(lin):
            d_a = d_a * (acts[i + 1] > 0)
        out[f"slot.to_k.{idx}.weight"] = d_a.reshape(-1, d).t() @ acts[i].reshape(-1, d)
        out[f"slot.to_k.{idx}.bias"] = d_a.reshape(-1, d).sum(0)
        d_a = d_a @ wl
    d_x = d_x + d_a                                                          # x + pe: pe is a constant
    d_pre = d_x * (x > 0)                                                    # conv1x1 + ReLU
    out["conv1x1.weight"] = (d_pre.reshape(-1, d).t() @ f.reshape(-1, ch)).reshape(d, ch, 1, 1)
    out["conv1x1.bias"] = d_pre.reshape(-1, d).sum(0)
    out["feat"] = (d_pre @ wc).permute(0, 2, 1).reshape(b, ch, h, w_)
    return out
