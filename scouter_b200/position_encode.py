"""Sine position encoding (reference ``sloter/utils/position_encode.py:10-46,77-87``).

The table depends only on (d, h, w); it is computed once per shape by ``scouter_pe_sine`` on the device
and cached token-major (n, d), the layout the fused head consumes.  ``forward`` keeps the reference's
return value, a (B, d, h, w) tensor (an expanded view of the cached table -- no per-call kernels).
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import _lib as L


class PositionEmbeddingSine(nn.Module):
    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        if not normalize or temperature != 10000 or (scale is not None and scale != 2 * math.pi):
            raise NotImplementedError("scouter_b200 implements the configuration SCOUTER builds: "
                                      "normalize=True, temperature=10000, scale=2*pi (position_encode.py:77-81)")
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        self.scale = 2 * math.pi
        self._cache = {}

    def table(self, h: int, w: int, device) -> torch.Tensor:
        """(h*w, d) fp32 on ``device``."""
        device = torch.device(device)
        if device.type != "cuda":
            raise L.ScouterError("PositionEmbeddingSine: CUDA only (no CPU path)")
        key = (h, w, device.index if device.index is not None else torch.cuda.current_device())
        t = self._cache.get(key)
        if t is None:
            d = 2 * self.num_pos_feats
            t = torch.empty(h * w, d, dtype=torch.float32, device=device)
            with torch.cuda.device(device):
                L.check(L.lib().scouter_pe_sine(t.data_ptr(), d, h, w, L.stream_ptr()), "scouter_pe_sine")
            self._cache[key] = t
        return t

    def forward(self, x):
        b, _, h, w = x.shape
        t = self.table(h, w, x.device)
        return t.view(h, w, -1).permute(2, 0, 1).unsqueeze(0).expand(b, -1, -1, -1).to(x.dtype)


def build_position_encoding(position_embedding, hidden_dim):
    n_steps = hidden_dim // 2
    if position_embedding in ("v2", "sine"):
        return PositionEmbeddingSine(n_steps, normalize=True)
    if position_embedding in ("v3", "learned"):
        raise NotImplementedError("learned position embedding is never built by SCOUTER (slot_model.py:73 uses 'sine')")
    raise ValueError(f"not supported {position_embedding}")
