"""xSlot attention module (reference ``sloter/utils/slot_attention.py:9-96``).

Same constructor signature, parameter names (``initial_slots``, ``to_q.0``, ``to_k.{0,2,..}``, ``gru.*``)
and return value ``(logits (B,C), loss)`` as the reference's ``SlotAttention``; ``ScouterAttention`` is an
alias (BASELINE.json's name for it).  ``to_q`` exists (checkpoint compatibility) but, as in the reference
(:52-53), is never applied.  The forward is one ``scouter_xslot_forward`` + ``scouter_head_finalize`` call.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
from torch import nn

from . import _lib as L


class SlotAttention(nn.Module):
    vis_dir = "sloter/vis"   # where the reference drops slot_{id}.png (:82-83)

    def __init__(self, num_classes, slots_per_class, dim, iters=3, eps=1e-8, vis=False, vis_id=0, loss_status=1,
                 power=1, to_k_layer=1):
        super().__init__()
        self.num_classes = num_classes
        self.slots_per_class = slots_per_class
        self.num_slots = num_classes * slots_per_class
        self.iters = iters
        self.eps = eps                       # stored, never used -- like the reference (:16)
        self.scale = dim ** -0.5
        self.loss_status = loss_status
        self.dim = dim

        # same RNG draw order as the reference constructor (:18-25); torch>=2 rejects a signed std in
        # torch.normal, torch 1.6 effectively computed mean + std * N(0,1)
        mu = torch.randn(1, 1, dim)
        sigma = torch.randn(1, 1, dim)
        sig = sigma.expand(1, self.num_slots, -1)
        self.initial_slots = nn.Parameter(mu.expand(1, self.num_slots, -1) + sig * torch.randn_like(sig))

        self.to_q = nn.Sequential(nn.Linear(dim, dim))
        layers = [nn.Linear(dim, dim)]
        for _ in range(1, to_k_layer):
            layers += [nn.ReLU(inplace=True), nn.Linear(dim, dim)]
        self.to_k = nn.Sequential(*layers)
        self.gru = nn.GRU(dim, dim)

        self.vis = vis
        self.vis_id = vis_id
        self.power = power
        self.keep_attn = False               # set True to retain the final attention in .last_attn
        self.last_attn = None
        self.last_vis_maps = None
        self._packed = None
        self._packed_sig = None
        self._desc = None

    # -- parameter pack -----------------------------------------------------------------------
    def invalidate(self):
        """Drop the packed parameter block (re-packed on the next forward); required after ``.data`` writes, which do not
        bump the version counter that ``desc_and_pack`` watches."""
        self._packed, self._packed_sig, self._desc = None, None, None

    def _linears(self):
        return [m for m in self.to_k if isinstance(m, nn.Linear)]

    def desc_and_pack(self, device):
        """(XSlotDesc, packed device buffer), re-packed when a parameter changed."""
        params = [self.initial_slots, self.gru.weight_ih_l0, self.gru.weight_hh_l0, self.gru.bias_ih_l0, self.gru.bias_hh_l0]
        lins = self._linears()
        for m in lins:
            params += [m.weight, m.bias]
        sig = tuple((p._version, p.data_ptr()) for p in params) + (self.loss_status, float(self.power), self.iters)
        if self._packed is not None and sig == self._packed_sig and self._packed.device == device:
            return self._desc, self._packed
        for p in params:
            if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                raise L.ScouterError(f"SlotAttention: parameters must be contiguous fp32 on {device} (found {p.dtype} on {p.device})")
        d = L.XSlotDesc()
        d.d = self.dim
        d.num_classes = self.num_classes
        d.slots_per_class = self.slots_per_class
        d.to_k_layers = len(lins)
        d.iters = self.iters
        d.loss_status = int(self.loss_status)
        d.power = float(self.power)
        d.initial_slots = self.initial_slots.data_ptr()
        for i, m in enumerate(lins):
            d.to_k_w[i] = m.weight.data_ptr()
            d.to_k_b[i] = m.bias.data_ptr()
        d.gru_w_ih = self.gru.weight_ih_l0.data_ptr()
        d.gru_w_hh = self.gru.weight_hh_l0.data_ptr()
        d.gru_b_ih = self.gru.bias_ih_l0.data_ptr()
        d.gru_b_hh = self.gru.bias_hh_l0.data_ptr()
        nbytes = L.lib().scouter_xslot_packed_bytes(C.byref(d))
        if nbytes == 0:
            L.check(-2, "scouter_xslot_packed_bytes")
        packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            L.check(L.lib().scouter_xslot_pack(C.byref(d), packed.data_ptr(), L.stream_ptr()), "scouter_xslot_pack")
        self._desc, self._packed, self._packed_sig = d, packed, sig
        return d, packed

    # -- vis branch (:68-85) ------------------------------------------------------------------
    def vis_maps(self, attn: torch.Tensor, vis_id=None, out_size=None, hw=None):
        """The explanation outputs on the device, all CUDA tensors:

        * ``maps`` (C,h,w) uint8 -- per-class sum of ``attn[vis_id]``, joint min-max, *255, truncation (:68-80);
        * ``heat`` (C,H,W) uint8 -- ``maps`` resized to ``out_size=(H,W)``, bit-identical to the reference's
          ``Image.open(slot_{id}.png).resize(image.size, Image.BILINEAR)`` (test.py:35); None without ``out_size``;
        * ``ratios`` (C) float64 -- ``sum(map)/(h*w*255)`` (test.py:43).
        """
        b, s, n = attn.shape
        vis_id = self.vis_id if vis_id is None else vis_id
        if hw is None:
            fs = int(n ** 0.5)                                   # the reference's square-map assumption (:76)
            hw = (fs, fs)
        if hw[0] * hw[1] != n:
            raise L.ScouterError(f"vis_maps: {n} tokens do not form a {hw[0]}x{hw[1]} map")
        if not attn.is_cuda or attn.dtype != torch.float32 or not attn.is_contiguous():
            raise L.ScouterError("vis_maps: attn must be a contiguous fp32 CUDA tensor (no CPU path)")
        dev = attn.device
        c = self.num_classes
        with torch.cuda.device(dev):
            maps = torch.empty(c, hw[0], hw[1], dtype=torch.uint8, device=dev)
            L.check(L.lib().scouter_vis_maps_u8(attn.data_ptr(), b, c, self.slots_per_class, n, vis_id, maps.data_ptr(),
                                                L.stream_ptr()), "scouter_vis_maps_u8")
            ratios = torch.empty(c, dtype=torch.float64, device=dev)
            heat = None
            oh = ow = 0
            if out_size is not None:
                oh, ow = int(out_size[0]), int(out_size[1])
                heat = torch.empty(c, oh, ow, dtype=torch.uint8, device=dev)
            L.check(L.lib().scouter_vis_upsample_u8(maps.data_ptr(), c, hw[0], hw[1], oh, ow, L.ptr(heat), ratios.data_ptr(),
                                                    L.stream_ptr()), "scouter_vis_upsample_u8")
        return maps, heat, ratios

    def emit_vis(self, attn: torch.Tensor, logits: torch.Tensor, hw=None):
        maps, _, _ = self.vis_maps(attn, hw=hw)
        self.last_vis_maps = maps.cpu().numpy()
        if self.vis_dir:
            from PIL import Image
            os.makedirs(self.vis_dir, exist_ok=True)
            for i, img in enumerate(self.last_vis_maps):
                Image.fromarray(img, mode="L").save(os.path.join(self.vis_dir, f"slot_{i:d}.png"))
        print(logits)

    def forward(self, inputs, inputs_x):
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("SlotAttention: training forward/backward is SURVEY.md row f1 (not implemented); "
                                      "use .eval() / torch.no_grad()")
        for t in (inputs, inputs_x):
            if not t.is_cuda or t.dtype != torch.float32:
                raise L.ScouterError("SlotAttention: inputs must be fp32 CUDA tensors (no CPU path)")
        b, n, d = inputs.shape
        dev = inputs.device
        with torch.cuda.device(dev):
            desc, packed = self.desc_and_pack(dev)
            want_attn = self.vis or self.keep_attn
            logits = torch.empty(b, self.num_classes, dtype=torch.float32, device=dev)
            attn = torch.empty(b, self.num_slots, n, dtype=torch.float32, device=dev) if want_attn else None
            attn_sum = torch.empty(b, dtype=torch.float32, device=dev)
            io = L.XSlotIO()
            io.batch, io.n = b, n
            io.x = inputs_x.data_ptr()
            io.x_sb, io.x_sn, io.x_sd = inputs_x.stride()
            io.x_pe = inputs.data_ptr()
            io.xpe_sb, io.xpe_sn, io.xpe_sd = inputs.stride()
            io.pe = 0
            io.logits, io.attn, io.attn_sum = logits.data_ptr(), L.ptr(attn), attn_sum.data_ptr()
            L.check(L.lib().scouter_xslot_forward(C.byref(desc), packed.data_ptr(), C.byref(io), 0, 0, L.stream_ptr()),
                    "scouter_xslot_forward")
            scratch = torch.empty(b, self.num_classes, dtype=torch.float32, device=dev)
            losses = torch.empty(3, dtype=torch.float32, device=dev)
            L.check(L.lib().scouter_head_finalize(logits.data_ptr(), attn_sum.data_ptr(), 0, b, self.num_classes,
                                                  self.num_slots, n, float(self.power), 0.0, scratch.data_ptr(),
                                                  losses.data_ptr(), L.stream_ptr()), "scouter_head_finalize")
            if self.keep_attn:
                self.last_attn = attn
            if self.vis:
                self.emit_vis(attn, logits)
        return logits, losses[2]


ScouterAttention = SlotAttention
