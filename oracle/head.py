"""Oracle (test infrastructure, see oracle/__init__.py): xSlot head on the CPU.

A plain-torch restatement of the reference's head, written from its math rather
than its code, parameterised by a ``state_dict`` so it needs no ``nn.Module``:

* ``sine_pe``          -- reference ``sloter/utils/position_encode.py:26-46`` as built by
                          ``build_position_encoding('sine', 64)`` (``:77-81``): 32 features per
                          axis, normalize=True, scale 2*pi, temperature 1e4.
* ``xslot_forward``    -- reference ``sloter/utils/slot_attention.py:44-96``.
* ``head_forward``     -- reference ``sloter/slot_model.py:108-125`` (conv1x1 + ReLU + PE + xSlot +
                          log_softmax [+ losses]).
* ``vis_maps_u8``      -- reference ``sloter/utils/slot_attention.py:68-80`` (the uint8 maps that
                          the vis branch writes as PNGs).

Pinned against the imported reference by ``oracle/make_golden.py`` /
``tests/test_oracle_golden.py`` (no reference-side tests exist: parity unpinned by them).
Every function takes ``dtype`` so the same restatement gives the fp64 "truth" used to
measure the reference's own fp32 noise floor.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def sine_pe(d: int, h: int, w: int, dtype=torch.float32) -> torch.Tensor:
    """(d, h, w) table; channels [0, d/2) encode the row, [d/2, d) the column.

    Follows position_encode.py:26-46: cumulative 1-based coordinates are divided by
    (last + 1e-6), scaled by 2*pi, divided by T^(2*floor(c/2)/(d/2)) and passed through
    sin (even c) / cos (odd c).  All in fp32 like the reference (then cast).
    """
    half = d // 2
    ys = torch.arange(1, h + 1, dtype=torch.float32)
    xs = torch.arange(1, w + 1, dtype=torch.float32)
    ys = ys / (ys[-1] + 1e-6) * (2 * math.pi)
    xs = xs / (xs[-1] + 1e-6) * (2 * math.pi)
    c = torch.arange(half, dtype=torch.float32)
    dim_t = 10000.0 ** (2 * torch.div(c, 2, rounding_mode="floor") / half)
    py = ys[:, None] / dim_t  # (h, half)
    px = xs[:, None] / dim_t  # (w, half)
    even = (torch.arange(half) % 2 == 0)
    py = torch.where(even, py.sin(), py.cos())
    px = torch.where(even, px.sin(), px.cos())
    out = torch.empty(d, h, w, dtype=torch.float32)
    out[:half] = py.t()[:, :, None].expand(half, h, w)
    out[half:] = px.t()[:, None, :].expand(half, h, w)
    return out.to(dtype)


def _gru_cell(u, s, w_ih, w_hh, b_ih, b_hh):
    """One PyTorch-convention GRU step, gate order [r | z | n] (nn.GRU docs)."""
    d = s.shape[-1]
    gi = u @ w_ih.t() + b_ih
    gh = s @ w_hh.t() + b_hh
    r = torch.sigmoid(gi[..., :d] + gh[..., :d])
    z = torch.sigmoid(gi[..., d:2 * d] + gh[..., d:2 * d])
    n = torch.tanh(gi[..., 2 * d:] + r * gh[..., 2 * d:])
    return (1 - z) * n + z * s


def xslot_forward(sd: dict, x_pe: torch.Tensor, x: torch.Tensor, *, num_classes: int,
                  slots_per_class: int, loss_status: int = 1, power: int = 1, iters: int = 3,
                  prefix: str = "", dtype=torch.float32, return_attn: bool = False):
    """slot_attention.py:44-96.  ``x_pe``/``x`` are (B, n, d); returns (logits (B,C), loss[, attn])."""
    g = lambda k: sd[prefix + k].to(dtype)
    x_pe, x = x_pe.to(dtype), x.to(dtype)
    b, n, d = x.shape
    k = x_pe
    li = 0
    while (prefix + f"to_k.{li}.weight") in sd:  # Linear at even indices, ReLU between (:30-37)
        if li:
            k = torch.relu(k)
        k = k @ g(f"to_k.{li}.weight").t() + g(f"to_k.{li}.bias")
        li += 2
    s = g("initial_slots").expand(b, -1, -1)
    w_ih, w_hh = g("gru.weight_ih_l0"), g("gru.weight_hh_l0")
    b_ih, b_hh = g("gru.bias_ih_l0"), g("gru.bias_hh_l0")
    scale = d ** -0.5
    for _ in range(iters):
        dots = torch.einsum("bid,bjd->bij", s, k) * scale                  # :55
        row = dots.sum(2, keepdim=True)                                    # r_bi
        tot = dots.sum(2).sum(1)[:, None, None]                            # t_b
        attn = torch.sigmoid(dots / row * tot)                             # :56-57 (no eps)
        upd = torch.einsum("bjd,bij->bid", x, attn) / d                    # :58-59 (divides by d)
        s = _gru_cell(upd, s, w_ih, w_hh, b_ih, b_hh)                      # :60-66
    if slots_per_class > 1:                                                # :87-91
        upd = upd.reshape(b, num_classes, slots_per_class, d).sum(2)
    logits = loss_status * upd.sum(2)                                      # :96
    loss = (attn.sum() / (b * attn.shape[1] * n)) ** power                 # :93-96 (relu(attn)==attn)
    if return_attn:
        return logits, loss, attn
    return logits, loss


def head_forward(sd: dict, feat: torch.Tensor, *, num_classes: int, slots_per_class: int,
                 loss_status: int = 1, power: int = 1, lambda_value: float = 1.0,
                 target: torch.Tensor | None = None, dtype=torch.float32, return_attn: bool = False):
    """slot_model.py:108-125.  ``feat`` is the backbone output viewed (B, ch, fs, fs) (NCHW)."""
    w = sd["conv1x1.weight"].to(dtype)
    bvec = sd["conv1x1.bias"].to(dtype)
    feat = feat.to(dtype)
    b, ch, h, wd = feat.shape
    x = torch.relu(F.conv2d(feat, w, bvec))
    d = x.shape[1]
    pe = sine_pe(d, h, wd, dtype)
    x_pe = (x + pe).reshape(b, d, -1).permute(0, 2, 1)
    x = x.reshape(b, d, -1).permute(0, 2, 1)
    logits, attn_loss, attn = xslot_forward(
        sd, x_pe, x, num_classes=num_classes, slots_per_class=slots_per_class, loss_status=loss_status,
        power=power, prefix="slot.", dtype=dtype, return_attn=True)
    out = F.log_softmax(logits, dim=1)
    res = {"log_probs": out, "logits": logits, "attn_loss": attn_loss}
    if return_attn:
        res["attn"] = attn
    if target is not None:
        nll = F.nll_loss(out, target)
        res["nll"] = nll
        res["loss"] = nll + lambda_value * attn_loss
    return res


def vis_maps_u8(attn: torch.Tensor, *, num_classes: int, slots_per_class: int, vis_id: int = 0):
    """slot_attention.py:68-80: per-class sum of the last attention, image ``vis_id``, joint min-max
    over (C, n), x255, truncate to uint8, reshape to (C, fs, fs)."""
    b, s, n = attn.shape
    a = attn
    if slots_per_class > 1:
        a = a.reshape(b, num_classes, slots_per_class, n).sum(2)
    a = a[vis_id]
    a = (a - a.min()) / (a.max() - a.min()) * 255.0
    fs = int(n ** 0.5)
    return a.reshape(a.shape[0], fs, fs).to(torch.float32).numpy().astype("uint8")


def preprocess_u8(images_u8_hwc, mean, std):
    """dataset/transform_func.py:63-66 (ToTensor: image/255 in float64, HWC->CHW) + :87-94 (Normalize) + engine.py:25
    (cast to float32).  ``images_u8_hwc``: (B,H,W,C) uint8 numpy array."""
    import numpy as np
    x = images_u8_hwc.astype(np.float64) / 255
    x = (x - np.asarray(mean, dtype=np.float64)) / np.asarray(std, dtype=np.float64)
    return torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2))).to(torch.float32)
