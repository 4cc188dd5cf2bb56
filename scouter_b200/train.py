"""Row f1: the training step of ``SlotModel`` on the device -- reference ``engine.py:28-35``
(``logits, loss_list = model(inputs, labels); loss.backward(); optimizer.step()``) and ``train.py:140-148`` (DDP, AdamW).

``SlotModel.forward`` in ``.train()`` mode runs through ``TrainEngine``:

* forward  -- the train-mode op program of ``plan.lower_backbone_train`` (conv and BatchNorm are separate ops: batch
  statistics cannot be folded into the weights; running statistics are updated in place like ``nn.BatchNorm2d``), every
  conv on the same tcgen05 / CUDA-core kernels as the eval path (``scouter_conv_forward``), then the fused head and
  ``scouter_head_finalize`` for the losses.  Every buffer stays alive for the backward.
* backward -- ``plan.backward_schedule`` executed op by op through the ``scouter_train_*`` entries of the C ABI (head backward,
  BatchNorm / conv / pool / split-attention backward); gradients of a block input that feeds both the main path and the
  shortcut are accumulated as the schedule's flags say.
* the step is exposed to autograd as ONE ``torch.autograd.Function`` whose inputs are the trainable parameters, so
  ``loss.backward()``, ``DistributedDataParallel`` (gradient all-reduce hooks fire on ``param.grad``, ``find_unused_parameters``
  covers ``slot.to_q``) and any ``torch.optim`` optimizer work unchanged on top of it -- which is what makes the module a
  drop-in for ``train.py``.  ``adamw_step`` is the fused alternative to ``torch.optim.AdamW`` over flat buffers.

Frozen layers (``dfs_freeze``, ``slot_model.py:79-93``): when no backbone parameter requires a gradient the backbone backward
is skipped altogether (the README's ``--freeze_layers 4`` recipe trains the head only); BatchNorm still runs in train mode,
as it does in the reference.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn.functional as F

from . import _lib as L
from .plan import backward_schedule, dgrad_weights, lower_backbone_train, split_weights_f16

POOL_KIND = {"maxpool": 0, "avgpool2": 1, "avgpool3": 2}


def _p(t):
    return 0 if t is None else t.data_ptr()


def _split_f16(w: torch.Tensor) -> torch.Tensor:
    """``plan.split_weights_f16`` in ONE launch (the step re-splits ~100 weight tensors; as torch ops that was ~750 launches)."""
    w = w.contiguous()
    out = torch.empty((2 * w.shape[0],) + tuple(w.shape[1:]), dtype=torch.int16, device=w.device)
    L.check(L.lib().scouter_split_weights_f16(w.data_ptr(), out.data_ptr(), w.numel(), L.stream_ptr()), "scouter_split_weights_f16")
    return out


class TrainEngine:
    def __init__(self, model):
        if not model.use_slot:
            raise NotImplementedError("training without the slot head (use_slot=False, the stage-1 classifier of README.md:30-36) "
                                      "is not implemented: scouter_b200 trains the xSlot model")
        self.model = model
        self.ops, self.nbuf, self.feat = lower_backbone_train(model.backbone)
        self.sched = backward_schedule(self.ops, self.feat)
        self._mods = dict(model.named_modules())
        self.saved = None

    # -- helpers ------------------------------------------------------------------------------------
    def _mod(self, key):
        return self._mods[key]

    def trainable(self):
        """(names, parameters) that require a gradient, in ``named_parameters`` order."""
        items = [(k, p) for k, p in self.model.named_parameters() if p.requires_grad]
        return [k for k, _ in items], [p for _, p in items]

    # -- forward ------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, target):
        m, lib, st = self.model, L.lib(), L.stream_ptr()
        math_mode = m._math()
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        B = x.shape[0]
        buf = {0: x.permute(0, 2, 3, 1).contiguous()}          # NHWC copy: the stem's weight gradient reads it
        sv = {"w": {}, "bn": {}, "att": {}}
        for i, op in enumerate(self.ops):
            k = op["kind"]
            src = buf[op["src"]]
            _, H, W, Cin = src.shape
            if k == "conv":
                conv = self._mod(op["key"])
                w = conv.weight.detach().permute(0, 2, 3, 1).contiguous()           # OHWI
                w2 = _split_f16(w) if math_mode == L.MATH_TC else None
                kh = conv.kernel_size[0]
                Ho = (H + 2 * op["pad"] - kh) // op["stride"] + 1
                Wo = (W + 2 * op["pad"] - kh) // op["stride"] + 1
                out = torch.empty(B, Ho, Wo, conv.out_channels, **f32)
                o = L.Op(kind=L.OP_CONV, src=0, src2=-1, dst=1, cin=Cin, cout=conv.out_channels, kh=kh, kw=kh, stride=op["stride"],
                         pad=op["pad"], groups=op["groups"], flags=0, mid=0, reserved=0, w=w.data_ptr(),
                         b=_p(None if conv.bias is None else conv.bias.detach()), w2=_p(w2), b2=0)
                if op["src"] == 0:      # the network's first conv: 1..4 input channels, NCHW in (resnet.py:401, slot_model.py:23-24)
                    L.check(lib.scouter_stem_conv_forward(C.byref(o), x.data_ptr(), out.data_ptr(), B, H, W, st), "scouter_stem_conv_forward")
                else:
                    L.check(lib.scouter_conv_forward(C.byref(o), src.data_ptr(), 0, out.data_ptr(), B, H, W, math_mode, st), "scouter_conv_forward")
                sv["w"][i] = w
                sv.setdefault("keep", []).append(w2)
                buf[op["dst"]] = out
            elif k == "bn":
                bn = self._mod(op["key"])
                Cc = src.shape[-1]
                y = torch.empty_like(src)
                mean, rstd = torch.empty(Cc, **f32), torch.empty(Cc, **f32)
                ws = torch.empty(2 * Cc, dtype=torch.float64, device=dev)
                ss = torch.empty(2 * Cc, **f32)
                res = buf[op["res"]] if op["res"] >= 0 else None
                a = L.BnTrainArgs(M=src.numel() // Cc, C=Cc, x=src.data_ptr(), sums=ws.data_ptr(), gamma=bn.weight.data_ptr(),
                                  beta=bn.bias.data_ptr(), running_mean=bn.running_mean.data_ptr(), running_var=bn.running_var.data_ptr(),
                                  scale=ss.data_ptr(), shift=ss.data_ptr() + 4 * Cc, save_mean=mean.data_ptr(), save_rstd=rstd.data_ptr(),
                                  eps=bn.eps, momentum=bn.momentum, residual=_p(res), y=y.data_ptr(), relu=int(op["relu"]))
                L.check(lib.scouter_train_bn_forward(C.byref(a), st), "scouter_train_bn_forward")
                bn.num_batches_tracked += 1
                sv["bn"][i] = (mean, rstd)
                sv.setdefault("keep", []).extend([ws, ss])
                buf[op["dst"]] = y
            elif k in POOL_KIND:
                kind = POOL_KIND[k]
                Ho, Wo = ((H + 1) // 2, (W + 1) // 2) if kind == 1 else ((H - 1) // 2 + 1, (W - 1) // 2 + 1)
                out = torch.empty(B, Ho, Wo, Cin, **f32)
                L.check(lib.scouter_pool_forward(kind, src.data_ptr(), out.data_ptr(), B, H, W, Cin, st), "scouter_pool_forward")
                buf[op["dst"]] = out
            elif k == "splat_gap":
                c = Cin // 2
                scratch = torch.empty(max(1, lib.scouter_splat_gap_scratch_floats(B, H * W, c)), **f32)
                gap = torch.empty(B, 1, 1, c, **f32)
                L.check(lib.scouter_splat_gap_forward(src.data_ptr(), scratch.data_ptr(), gap.data_ptr(), B, H * W, c, st),
                        "scouter_splat_gap_forward")
                sv.setdefault("keep", []).append(scratch)
                buf[op["dst"]] = gap
            elif k == "splat_mix":
                c = Cin // 2
                lg = buf[op["logits"]]
                out = torch.empty(B, H, W, c, **f32)
                L.check(lib.scouter_splat_apply_forward(src.data_ptr(), lg.data_ptr(), out.data_ptr(), B, H, W, c, st),
                        "scouter_splat_apply_forward")
                sv["att"][op["dst"]] = torch.softmax(lg.reshape(B, 2, c), dim=1).contiguous()
                buf[op["dst"]] = out
            else:
                raise L.ScouterError(f"train forward: unknown op kind {k}")
        # ---- head (slot_model.py:108-125) -----------------------------------------------------------------
        feat = buf[self.feat]
        _, fh, fw, ch = feat.shape
        if ch != m.channel:
            raise L.ScouterError(f"SlotModel: backbone produces {ch} channels but args.channel={m.channel}")
        n, s, c = fh * fw, m.slot.num_slots, m.slot.num_classes
        logits, log_probs = torch.empty(B, c, **f32), torch.empty(B, c, **f32)
        attn_sum, losses = torch.empty(B, **f32), torch.zeros(3, **f32)
        pe = m.position_emb.table(fh, fw, dev)
        io = L.HeadIO()
        io.batch, io.h, io.w, io.channel, io.layout, io.math = B, fh, fw, ch, L.LAYOUT_NHWC, math_mode
        io.feat, io.pe = feat.data_ptr(), pe.data_ptr()
        io.conv_w, io.conv_b = m.conv1x1.weight.data_ptr(), m.conv1x1.bias.data_ptr()
        io.logits, io.attn, io.attn_sum, io.x_out, io.conv_w_split = logits.data_ptr(), 0, attn_sum.data_ptr(), 0, 0
        desc, packed = m.slot.desc_and_pack(dev)
        nbytes = lib.scouter_head_workspace_bytes(C.byref(desc), C.byref(io))
        ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 1024
        L.check(lib.scouter_head_forward(C.byref(desc), packed.data_ptr(), C.byref(io), ws.data_ptr() + off, nbytes, st), "scouter_head_forward")
        L.check(lib.scouter_head_finalize(logits.data_ptr(), attn_sum.data_ptr(), _p(target), B, c, s, n, float(m.slot.power),
                                          m.lambda_value, log_probs.data_ptr(), losses.data_ptr(), st), "scouter_head_finalize")
        m.feature_size = fh
        m._sig = None            # the BatchNorm kernels updated running statistics in place: eval-mode folded weights are stale
        self.saved = dict(buf=buf, sv=sv, feat=feat, fhw=(fh, fw), log_probs=log_probs, attn_sum=attn_sum, target=target, pe=pe)
        return log_probs, losses

    # -- backward -----------------------------------------------------------------------------------
    @torch.no_grad()
    def backward(self, g_nll, g_attn):
        """Gradients of ``g_nll * nll + g_attn * attn_loss`` w.r.t. every parameter -> {state_dict key: tensor}."""
        m, lib, st = self.model, L.lib(), L.stream_ptr()
        S_ = self.saved
        if S_ is None or S_["target"] is None:
            raise L.ScouterError("train backward: no saved forward with targets")
        buf, sv, feat = S_["buf"], S_["sv"], S_["feat"]
        dev = feat.device
        f32 = dict(dtype=torch.float32, device=dev)
        B, fh, fw, ch = feat.shape
        n, s, c, spc = fh * fw, m.slot.num_slots, m.slot.num_classes, m.slot.slots_per_class
        lins = m.slot._linears()
        nl, iters = len(lins), m.slot.iters
        # d nll / d logits through log_softmax + nll_loss (mean over the batch); attention-area term (slot_attention.py:93-96)
        g_logits = ((S_["log_probs"].exp() - F.one_hot(S_["target"], c).to(torch.float32)) * (float(g_nll) / B)).contiguous()
        cnt = float(B * s * n)
        mean_attn = S_["attn_sum"].double().sum() / cnt
        power = float(m.slot.power)
        coef = (float(g_attn) * power * mean_attn ** (power - 1.0) / cnt).to(torch.float32).reshape(1).contiguous()
        grads = {}
        z = lambda p: torch.zeros(p.shape, **f32)
        need_backbone = any(p.requires_grad for p in m.backbone.parameters())
        # the two ch-sized products of the head backward (d feat = d pre . W and d W = d pre^T . feat, ch = 2048) leave the
        # one-CTA-per-image kernel: it hands d pre (B, n, 64) over and the GEMMs run on the conv kernels below
        d_pre = torch.empty(B, n, 64, **f32)
        g = {"conv1x1.weight": z(m.conv1x1.weight), "conv1x1.bias": z(m.conv1x1.bias),
             "slot.gru.weight_ih_l0": z(m.slot.gru.weight_ih_l0), "slot.gru.weight_hh_l0": z(m.slot.gru.weight_hh_l0),
             "slot.gru.bias_ih_l0": z(m.slot.gru.bias_ih_l0), "slot.gru.bias_hh_l0": z(m.slot.gru.bias_hh_l0),
             "slot.initial_slots": z(m.slot.initial_slots)}
        a = L.HeadBwdArgs(B=B, n=n, ch=ch, S=s, C=c, spc=spc, L=nl, iters=iters, loss_status=int(m.slot.loss_status),
                          feat=feat.data_ptr(), conv_w=m.conv1x1.weight.data_ptr(), conv_b=m.conv1x1.bias.data_ptr(), pe=S_["pe"].data_ptr(),
                          w_ih=m.slot.gru.weight_ih_l0.data_ptr(), w_hh=m.slot.gru.weight_hh_l0.data_ptr(),
                          b_ih=m.slot.gru.bias_ih_l0.data_ptr(), b_hh=m.slot.gru.bias_hh_l0.data_ptr(),
                          slots0=m.slot.initial_slots.data_ptr(), g_logits=g_logits.data_ptr(), attn_coef=coef.data_ptr(),
                          d_feat=0, d_pre=d_pre.data_ptr(), g_conv_w=g["conv1x1.weight"].data_ptr(), g_conv_b=g["conv1x1.bias"].data_ptr(),
                          g_w_ih=g["slot.gru.weight_ih_l0"].data_ptr(), g_w_hh=g["slot.gru.weight_hh_l0"].data_ptr(),
                          g_b_ih=g["slot.gru.bias_ih_l0"].data_ptr(), g_b_hh=g["slot.gru.bias_hh_l0"].data_ptr(),
                          g_slots0=g["slot.initial_slots"].data_ptr())
        for l, lin in enumerate(lins):
            kw, kb = f"slot.to_k.{2 * l}.weight", f"slot.to_k.{2 * l}.bias"
            g[kw], g[kb] = z(lin.weight), z(lin.bias)
            a.to_k_w[l], a.to_k_b[l], a.g_to_k_w[l], a.g_to_k_b[l] = lin.weight.data_ptr(), lin.bias.data_ptr(), g[kw].data_ptr(), g[kb].data_ptr()
        per = lib.scouter_train_head_backward_scratch_floats(n, s, nl, iters)
        scratch = torch.empty(B * per, **f32)
        a.scratch, a.scratch_per_image = scratch.data_ptr(), per
        L.check(lib.scouter_train_head_backward(C.byref(a), st), "scouter_train_head_backward")
        # d conv1x1.weight (64, ch) = d pre^T . feat: the weight gradient of a 1x1 conv ch -> 64 on the (B, fh, fw) map
        wa = L.WgradArgs(B=B, H=fh, W=fw, Cin=ch, Ho=fh, Wo=fw, Cout=64, k=1, stride=1, pad=0, groups=1, x=feat.data_ptr(),
                         dy=d_pre.data_ptr(), dw=g["conv1x1.weight"].data_ptr(), db=0)
        L.check(lib.scouter_train_conv_wgrad(C.byref(wa), st), "scouter_train_conv_wgrad(conv1x1)")
        grads.update(g)
        if not need_backbone:
            return grads
        # d feat (B, n, ch) = d pre . W: a 1x1 conv 64 -> ch with the transposed weights, on the forward tcgen05 kernel
        math_mode = m._math()
        wt = m.conv1x1.weight.detach().reshape(64, ch).t().contiguous()                # (ch, 64) = OHWI (ch, 1, 1, 64)
        wt2 = _split_f16(wt) if math_mode == L.MATH_TC else None
        d_feat = torch.empty(B, n, ch, **f32)
        op = L.Op(kind=L.OP_CONV, src=0, src2=-1, dst=1, cin=64, cout=ch, kh=1, kw=1, stride=1, pad=0, groups=1, flags=0, mid=0,
                  reserved=0, w=wt.data_ptr(), b=0, w2=_p(wt2), b2=0)
        L.check(lib.scouter_conv_forward(C.byref(op), d_pre.data_ptr(), 0, d_feat.data_ptr(), B, fh, fw, math_mode, st),
                "scouter_conv_forward(d feat)")
        # ---- backbone: the reverse schedule ------------------------------------------------------------------
        d = {self.feat: d_feat.reshape(B, fh, fw, ch)}

        def put(e, b, gt):
            d[b] = d[b] + gt if e["acc"][b] else gt

        bn_index = {op["dst"]: i for i, op in enumerate(self.ops) if op["kind"] == "bn"}
        conv_index = {op["dst"]: i for i, op in enumerate(self.ops) if op["kind"] == "conv"}
        for e in self.sched:
            k = e["kind"]
            if k == "conv":
                conv = self._mod(e["key"])
                x, dy = buf[e["src"]], d[e["dst"]].contiguous()
                w = sv["w"][conv_index[e["dst"]]]
                _, H, W, Cin = x.shape
                _, Ho, Wo, Cout = dy.shape
                dw = torch.zeros_like(w)
                db = None if conv.bias is None else torch.zeros(Cout, **f32)
                wa = L.WgradArgs(B=B, H=H, W=W, Cin=Cin, Ho=Ho, Wo=Wo, Cout=Cout, k=w.shape[1], stride=e["stride"], pad=e["pad"],
                                 groups=e["groups"], x=x.data_ptr(), dy=dy.data_ptr(), dw=dw.data_ptr(), db=_p(db))
                L.check(lib.scouter_train_conv_wgrad(C.byref(wa), st), "scouter_train_conv_wgrad")
                grads[e["key"] + ".weight"] = dw.permute(0, 3, 1, 2).contiguous()
                if db is not None:
                    grads[e["key"] + ".bias"] = db
                if e["need_dx"]:
                    dx = torch.empty_like(x)
                    kk = w.shape[1]
                    if e["stride"] == 1 and (Cout // e["groups"]) % 16 == 0:
                        # the data gradient of a stride-1 conv IS a conv (taps flipped, channels swapped inside each group,
                        # padding k-1-pad): it runs on the forward kernels -- tcgen05 in the tensor-core mode
                        math_mode = m._math()
                        wt = dgrad_weights(w, e["groups"])
                        wt2 = _split_f16(wt) if math_mode == L.MATH_TC else None
                        o = L.Op(kind=L.OP_CONV, src=0, src2=-1, dst=1, cin=Cout, cout=Cin, kh=kk, kw=kk, stride=1, pad=kk - 1 - e["pad"],
                                 groups=e["groups"], flags=0, mid=0, reserved=0, w=wt.data_ptr(), b=0, w2=_p(wt2), b2=0)
                        L.check(lib.scouter_conv_forward(C.byref(o), dy.data_ptr(), 0, dx.data_ptr(), B, Ho, Wo, math_mode, st), "scouter_conv_forward (dgrad)")
                    else:
                        da = L.DgradArgs(B=B, H=H, W=W, Cin=Cin, Ho=Ho, Wo=Wo, Cout=Cout, k=kk, stride=e["stride"], pad=e["pad"],
                                         groups=e["groups"], dy=dy.data_ptr(), w=w.data_ptr(), dx=dx.data_ptr())
                        L.check(lib.scouter_train_conv_dgrad(C.byref(da), st), "scouter_train_conv_dgrad")
                    put(e, e["src"], dx)
            elif k == "bn":
                bn = self._mod(e["key"])
                x, out, dy = buf[e["src"]], buf[e["dst"]], d[e["dst"]].contiguous()
                mean, rstd = sv["bn"][bn_index[e["dst"]]]
                Cc = x.shape[-1]
                dg, dbt = torch.zeros(Cc, **f32), torch.zeros(Cc, **f32)
                ws = torch.empty(2 * Cc, dtype=torch.float64, device=dev)
                coef_ = torch.empty(3 * Cc, **f32)
                dx = torch.empty_like(x)
                dres = torch.empty_like(x) if e["res"] >= 0 else None
                ba = L.BnBwdArgs(M=x.numel() // Cc, C=Cc, x=x.data_ptr(), out=out.data_ptr(), d_out=dy.data_ptr(), gamma=bn.weight.data_ptr(),
                                 save_mean=mean.data_ptr(), save_rstd=rstd.data_ptr(), sums=ws.data_ptr(), d_gamma=dg.data_ptr(),
                                 d_beta=dbt.data_ptr(), coef=coef_.data_ptr(), dx=dx.data_ptr(), d_residual=_p(dres), relu=int(e["relu"]))
                L.check(lib.scouter_train_bn_backward(C.byref(ba), st), "scouter_train_bn_backward")
                grads[e["key"] + ".weight"], grads[e["key"] + ".bias"] = dg, dbt
                put(e, e["src"], dx)
                if e["res"] >= 0:
                    put(e, e["res"], dres)
            elif k in POOL_KIND:
                x, dy = buf[e["src"]], d[e["dst"]].contiguous()
                dx = torch.empty_like(x)
                pa = L.PoolBwdArgs(B=B, H=x.shape[1], W=x.shape[2], C=x.shape[3], Ho=dy.shape[1], Wo=dy.shape[2], x=x.data_ptr(),
                                   dy=dy.data_ptr(), dx=dx.data_ptr())
                L.check(lib.scouter_train_pool_backward(C.byref(pa), POOL_KIND[k], st), "scouter_train_pool_backward")
                put(e, e["src"], dx)
            elif k == "splat_mix":
                x2, dy, att = buf[e["src"]], d[e["dst"]].contiguous(), sv["att"][e["dst"]]
                cc = att.shape[2]
                d_att, d_logit = torch.empty_like(att), torch.empty_like(att)
                sa = L.SplatBwdArgs(B=B, HW=x2.shape[1] * x2.shape[2], C=cc, x2=x2.data_ptr(), d_out=dy.data_ptr(), att=att.data_ptr(),
                                    d_att=d_att.data_ptr(), d_logit=d_logit.data_ptr(), d_gap=0, d_x2=0)
                L.check(lib.scouter_train_splat_backward(C.byref(sa), 0, st), "scouter_train_splat_backward")
                put(e, e["logits"], d_logit.reshape(B, 1, 1, 2 * cc))
            elif k == "splat_gap":
                x2, dmix, att = buf[e["src"]], d[e["d_mix"]].contiguous(), sv["att"][e["d_mix"]]
                cc = att.shape[2]
                d_gap = d[e["dst"]].reshape(B, cc).contiguous()
                d_x2 = torch.empty_like(x2)
                sa = L.SplatBwdArgs(B=B, HW=x2.shape[1] * x2.shape[2], C=cc, x2=x2.data_ptr(), d_out=dmix.data_ptr(), att=att.data_ptr(),
                                    d_att=0, d_logit=0, d_gap=d_gap.data_ptr(), d_x2=d_x2.data_ptr())
                L.check(lib.scouter_train_splat_backward(C.byref(sa), 1, st), "scouter_train_splat_backward")
                put(e, e["src"], d_x2)
            else:
                raise L.ScouterError(f"train backward: unknown op kind {k}")
        return grads


class _TrainStepFn(torch.autograd.Function):
    """(x, target, *trainable parameters) -> (log_probs, loss, nll, attn_loss); the parameters are inputs only so that
    autograd routes their gradients (and DDP's hooks see them) -- the arithmetic reads the live tensors."""

    @staticmethod
    def forward(ctx, engine, x, target, *params):
        log_probs, losses = engine.forward(x, target)
        ctx.engine = engine
        ctx.n_params = len(params)
        ctx.mark_non_differentiable(log_probs)
        return log_probs, losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, _g_out, g_loss, g_nll, g_attn):
        eng = ctx.engine
        lam = eng.model.lambda_value
        gl = 0.0 if g_loss is None else float(g_loss)
        gn = gl + (0.0 if g_nll is None else float(g_nll))                  # loss = nll + lambda * attn_loss (slot_model.py:122)
        ga = gl * lam + (0.0 if g_attn is None else float(g_attn))
        grads = eng.backward(gn, ga)
        names, params = eng.trainable()
        out = []
        for k, p in zip(names, params):
            g = grads.get(k)
            out.append(None if g is None else g.reshape(p.shape))          # slot.to_q.* never gets one
        eng.saved = None
        return (None, None, None, *out)


def train_forward(model, x, target):
    """``SlotModel.forward`` in train mode: ``[log_probs, [loss, nll, attn_loss]]`` with autograd edges to the parameters."""
    eng = model.__dict__.get("_train_engine")
    if eng is None:
        eng = TrainEngine(model)
        model.__dict__["_train_engine"] = eng
    if target is None:
        with torch.no_grad():
            tgt = torch.zeros(x.shape[0], dtype=torch.int64, device=x.device)
            log_probs, _ = eng.forward(x, tgt)
        eng.saved = None
        return log_probs
    _, params = eng.trainable()
    if torch.is_grad_enabled() and params:
        out, loss, nll, attn_loss = _TrainStepFn.apply(eng, x, target, *params)
    else:
        out, ls = eng.forward(x, target)
        eng.saved = None
        loss, nll, attn_loss = ls[0], ls[1], ls[2]
    return [out, [loss, nll, attn_loss]]


def adamw_step(params, grads, exp_avg, exp_avg_sq, lr, step, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    """In-place fused AdamW (``torch.optim.AdamW`` single-tensor semantics, train.py:146) over flat fp32 CUDA tensors of equal
    length; ``step`` is the 1-based step count.  Covers only what is passed in: leave frozen / gradient-less parameters out
    (torch skips ``grad is None`` -- they must not be decayed)."""
    for t in (params, grads, exp_avg, exp_avg_sq):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != params.numel():
            raise L.ScouterError("adamw_step: flat contiguous fp32 CUDA tensors of one length expected")
    b1, b2 = betas
    a = L.AdamWArgs(decay=1 - lr * weight_decay, one_minus_beta1=1 - b1, beta2=b2, one_minus_beta2=1 - b2, eps=eps,
                    step_size=lr / (1 - b1 ** step), bias_correction2_sqrt=math.sqrt(1 - b2 ** step), n=params.numel(),
                    p=params.data_ptr(), g=grads.data_ptr(), m=exp_avg.data_ptr(), v=exp_avg_sq.data_ptr())
    with torch.cuda.device(params.device):
        L.check(L.lib().scouter_train_adamw_step(C.byref(a), L.stream_ptr()), "scouter_train_adamw_step")
