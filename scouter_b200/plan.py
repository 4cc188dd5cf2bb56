"""Lowering of the module tree to the library's op program, and the runners that execute it.

* ``lower_backbone``   walks a ``backbone.ResNet`` (mirror of timm/models/resnet.py:491-509,
                       resnest.py:111-143, split_attn.py:54-80), folds eval-mode BatchNorm into every conv
                        and emits ``scouter_op_t`` records + the packed OHWI weights.
* ``BackboneRunner``   ``backbone(x)`` standalone: same return value as the reference's
                       ``ResNet.forward`` (classifier output, or the NCHW-flattened features once
                       pool/fc are ``Identical``).
* ``CompiledProgram``  a ``scouter_plan_t`` handle bound to one input shape and its torch-allocated arena;
                       ``slot_model.SlotModel`` runs it followed by the fused xSlot head + finalize
                       (slot_model.py:105-127), optionally inside a CUDA graph.

All device memory (arena, workspaces, outputs) is torch-allocated and cached per input shape; the
library only sees raw pointers.  Weight packs are rebuilt when any parameter/buffer version changes.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
from torch import nn

from . import _lib as L


def _require_cuda(x: torch.Tensor, what: str):
    if not x.is_cuda:
        raise L.ScouterError(f"{what}: input is on {x.device}; scouter_b200 runs on CUDA (sm_100a) only, there is no CPU path")
    if x.dtype != torch.float32:
        raise L.ScouterError(f"{what}: input dtype {x.dtype}; the reference casts to float32 (engine.py:25) and so must the caller")


def _version_signature(module: nn.Module):
    """Changes when a parameter / buffer is re-assigned or modified in place through autograd-visible ops.  In-place
    writes through ``.data`` (``p.data.copy_()``, ``bn.running_mean.data...``) do NOT bump ``_version``: after those, call
    ``SlotModel.invalidate()`` / ``BackboneRunner.invalidate()``."""
    sig = 0
    for t in list(module.parameters()) + list(module.buffers()):
        sig = (sig * 1000003 + t._version * 31 + t.data_ptr()) & 0xFFFFFFFFFFFFFFFF
    return sig


def fold_conv_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d | None):
    """(Cout, kh, kw, Cin/g) fp32 OHWI weight and (Cout) bias with eval-mode BN folded (fp64 fold, fp32 store)."""
    w = conv.weight.detach().double()
    b = conv.bias.detach().double() if conv.bias is not None else torch.zeros(w.shape[0], dtype=torch.float64, device=w.device)
    if bn is not None:
        s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
        w = w * s[:, None, None, None]
        b = (b - bn.running_mean.detach().double()) * s + bn.bias.detach().double()
    return w.permute(0, 2, 3, 1).contiguous().float(), b.float().contiguous()


def round_tf32(t: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest tf32 (ties away, like cvt.rna.tf32.f32), kept in fp32 storage."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def trunc19_remainder(t: torch.Tensor) -> torch.Tensor:
    """x - trunc19(x): what is left after kind::tf32 reads the top 19 bits of an fp32 word (exact in fp32)."""
    hi = (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
    return t - hi


def split_weights_bf16(w: torch.Tensor) -> torch.Tensor:
    """(2*Cout, kh, kw, Cin/g) bf16: rows [0,Cout) = bf16(W), rows [Cout,2Cout) = bf16(W - trunc19(W))."""
    return torch.cat([w, trunc19_remainder(w)], dim=0).to(torch.bfloat16).contiguous()


def split_weights_f16(w: torch.Tensor) -> torch.Tensor:
    """(2*Cout, kh, kw, Cin/g) 16-bit words (int16 view): rows [0,Cout) = fp16(W) (round to nearest, saturating at +-65504),
    rows [Cout,2Cout) = bf16(W - fp16(W)) -- the pre-split weight operand of the error-compensated tcgen05 convs
    (csrc/ptx.cuh split2_wgt, umma_conv.cu, umma_halo.cu; ``w2`` of ``scouter_op_t``)."""
    w = w.contiguous()
    h = w.clamp(-65504.0, 65504.0).to(torch.float16)
    r = (w - h.float()).to(torch.bfloat16)
    return torch.cat([h.view(torch.int16), r.view(torch.int16)], dim=0).contiguous()


def dgrad_weights(w_ohwi: torch.Tensor, groups: int = 1) -> torch.Tensor:
    """Row f1 groundwork (host side; no backward program exists yet): the data gradient of a stride-1 convolution is
    itself a convolution -- ``dX = conv(dY, W')`` with the taps flipped, input/output channels swapped inside each
    group and padding ``k - 1 - pad`` -- so it can run on the forward kernels unchanged.  ``w_ohwi`` is the forward
    weight in this library's layout (Cout, kh, kw, Cin/g); returns W' as (Cin, kh, kw, Cout/g)."""
    cout, kh, kw, cin_g = w_ohwi.shape
    if cout % groups:
        raise L.ScouterError(f"dgrad_weights: {cout} output channels do not split into {groups} groups")
    w = w_ohwi.reshape(groups, cout // groups, kh, kw, cin_g).flip(2, 3)          # flip the taps
    return w.permute(0, 4, 2, 3, 1).reshape(groups * cin_g, kh, kw, cout // groups).contiguous()


class Program:
    """Accumulates ops; keeps every packed tensor alive."""

    def __init__(self, math: int = L.MATH_FP32, fast_stages=()):
        self.math = math
        self.fast_stages = frozenset(fast_stages)   # per-stage precision policy (MATH_TC only): stages run as single-pass tf32
        self.fast = False                           # set by stage(): ops emitted now carry F_TF32_1PASS
        self.ops: list[L.Op] = []
        self.keep: list[torch.Tensor] = []
        self.nbuf = 1  # buffer 0 = network input

    def buf(self) -> int:
        self.nbuf += 1
        return self.nbuf - 1

    def stage(self, name: str):
        self.fast = self.math == L.MATH_TC and name in self.fast_stages

    def _t(self, t):
        if t is None:
            return 0
        self.keep.append(t)
        return t.data_ptr()

    def emit(self, kind, src, dst, *, src2=-1, cin=0, cout=0, k=1, stride=1, pad=0, groups=1, flags=0, mid=0,
             w=None, b=None, w2=None, b2=None):
        if self.fast:
            flags |= L.F_TF32_1PASS
        self.ops.append(L.Op(kind=kind, src=src, src2=src2, dst=dst, cin=cin, cout=cout, kh=k, kw=k, stride=stride,
                             pad=pad, groups=groups, flags=flags, mid=mid, reserved=0,
                             w=self._t(w), b=self._t(b), w2=self._t(w2), b2=self._t(b2)))
        return dst

    def conv(self, src, conv: nn.Conv2d, bn, *, relu, residual=-1, stem=False):
        w, b = fold_conv_bn(conv, bn)
        if (self.math == L.MATH_TC_FAST or self.fast) and not stem:
            w = round_tf32(w)        # 1-pass kind::tf32 reads the top 19 bits: make that a rounding, not a truncation
        w2 = None
        if self.math == L.MATH_TC and not self.fast and not stem:
            # pre-split operand of the error-compensated kernels: [fp16 W ; bf16 (W - fp16 W)]
            w2 = split_weights_f16(w)
        flags = (L.F_RELU if relu else 0) | (L.F_RESIDUAL if residual >= 0 else 0)
        return self.emit(L.OP_STEM_CONV if stem else L.OP_CONV, src, self.buf(), src2=residual,
                         cin=conv.in_channels, cout=conv.out_channels, k=conv.kernel_size[0], stride=conv.stride[0],
                         pad=conv.padding[0], groups=conv.groups, flags=flags, w=w, b=b, w2=w2)


def _lower_resnest_block(p: Program, blk, x: int) -> int:
    t = p.conv(x, blk.conv1, blk.bn1, relu=True)
    sa = blk.conv2
    c = sa.conv.out_channels // 2
    t2 = p.conv(t, sa.conv, sa.bn0, relu=True)                                   # (B,H,W,2C)
    gap = p.emit(L.OP_SPLAT_GAP, t2, p.buf(), cout=c)                            # (B,1,1,C)
    hid = p.conv(gap, sa.fc1, sa.bn1, relu=True)                                 # fc1 + bn1 + ReLU  (B,1,1,mid)
    av = p.conv(hid, sa.fc2, None, relu=False)                                   # fc2 logits        (B,1,1,2C)
    t3 = p.emit(L.OP_SPLAT_APPLY, t2, p.buf(), src2=av, cout=c, flags=L.F_AVD_POOL if blk.avd_last is not None else 0)
    res = x
    if blk.downsample is not None:
        pool, dconv, dbn = blk.downsample[0], blk.downsample[1], blk.downsample[2]
        r = x
        if isinstance(pool, nn.AvgPool2d):
            r = p.emit(L.OP_AVGPOOL, x, p.buf(), k=2, stride=pool.stride if isinstance(pool.stride, int) else pool.stride[0],
                       pad=0, flags=L.F_CEIL_MODE)
        res = p.conv(r, dconv, dbn, relu=False)
    return p.conv(t3, blk.conv3, blk.bn3, relu=True, residual=res)


def _lower_basic_block(p: Program, blk, x: int) -> int:
    t = p.conv(x, blk.conv1, blk.bn1, relu=True)
    res = x
    if blk.downsample is not None:
        res = p.conv(x, blk.downsample[0], blk.downsample[1], relu=False)
    return p.conv(t, blk.conv2, blk.bn2, relu=True, residual=res)


STAGES = ("stem", "layer1", "layer2", "layer3", "layer4")


def default_fast_stages():
    """SCOUTER_TC_FAST_STAGES=stem,layer1,...: the stages of a SCOUTER_MATH_TC backbone that run as ONE tf32 pass on rounded
    operands instead of the error-compensated product (the precision-policy experiment of profiles/r02_precision_policy.txt).
    Default: none -- no policy passes the parity bars of all configurations with margin."""
    v = os.environ.get("SCOUTER_TC_FAST_STAGES", "").strip().lower()
    st = tuple(t for t in (x.strip() for x in v.split(",")) if t)
    bad = [t for t in st if t not in STAGES]
    if bad:
        raise L.ScouterError(f"SCOUTER_TC_FAST_STAGES: unknown stage(s) {bad}; expected a subset of {list(STAGES)}")
    return st


def lower_backbone(net, math: int = L.MATH_FP32, fast_stages=()):
    """Returns (program, feature_buffer_id) for ``forward_features`` (resnet.py:491-501)."""
    from .backbone import BasicBlock, ResNestBottleneck
    p = Program(math, fast_stages)
    p.stage("stem")
    if isinstance(net.conv1, nn.Sequential):                                    # deep stem
        x = p.conv(0, net.conv1[0], net.conv1[1], relu=True, stem=True)
        x = p.conv(x, net.conv1[3], net.conv1[4], relu=True)
        x = p.conv(x, net.conv1[6], net.bn1, relu=True)
    else:
        x = p.conv(0, net.conv1, net.bn1, relu=True, stem=True)
    x = p.emit(L.OP_MAXPOOL, x, p.buf(), k=3, stride=2, pad=1)
    for li in range(1, 5):
        p.stage(f"layer{li}")
        for blk in getattr(net, f"layer{li}"):
            if isinstance(blk, ResNestBottleneck):
                x = _lower_resnest_block(p, blk, x)
            elif isinstance(blk, BasicBlock):
                x = _lower_basic_block(p, blk, x)
            else:
                raise L.ScouterError(f"cannot lower block type {type(blk).__name__}")
    p.stage("")
    return p, x


# ---------------------------------------------------------------------------------------------------------------------
# Row f1 groundwork: the TRAINING step as an op program (host side only -- the C++ executor does not run these kinds yet;
# tests/test_train_program.py interprets the program with the host emulations of the draft kernels in csrc/draft/ and
# compares every gradient with the train-mode oracle).  In train mode BatchNorm uses batch statistics and cannot be
# folded into the conv, so conv and BatchNorm(+residual, ReLU) are separate ops and every buffer that the backward needs
# (conv inputs, BatchNorm inputs and outputs, pool inputs) stays live until its backward op has run.
# ---------------------------------------------------------------------------------------------------------------------
def lower_backbone_train(net, prefix: str = "backbone."):
    """Forward op list of ``ResNet.forward_features`` (resnet.py:491-501) in train mode.  Ops are dicts:
    ``conv`` {src, dst, key, stride, pad, groups, need_dx}, ``bn`` {src, dst, key, relu, res}, ``maxpool`` /
    ``avgpool2`` (2,2,ceil,count_include_pad=False) / ``avgpool3`` (3,2,1) {src, dst}, ``splat_gap`` {src, dst} and
    ``splat_mix`` {src (x2), logits, dst} of split attention (split_attn.py:62-79).  Returns (ops, n_buffers, feature
    buffer); buffer 0 is the network input, keys are the reference's state_dict prefixes."""
    from .backbone import BasicBlock, ResNestBottleneck
    ops, nbuf = [], [1]

    def buf():
        nbuf[0] += 1
        return nbuf[0] - 1

    def conv_bn(src, ckey, bkey, conv, relu=True, res=-1, need_dx=True):
        mid = buf()
        ops.append(dict(kind="conv", src=src, dst=mid, key=prefix + ckey, stride=conv.stride[0], pad=conv.padding[0],
                        groups=conv.groups, need_dx=need_dx))
        dst = buf()
        ops.append(dict(kind="bn", src=mid, dst=dst, key=prefix + bkey, relu=relu, res=res))
        return dst

    if isinstance(net.conv1, nn.Sequential):
        x = conv_bn(0, "conv1.0", "conv1.1", net.conv1[0], need_dx=False)
        x = conv_bn(x, "conv1.3", "conv1.4", net.conv1[3])
        x = conv_bn(x, "conv1.6", "bn1", net.conv1[6])
    else:
        x = conv_bn(0, "conv1", "bn1", net.conv1, need_dx=False)
    dst = buf()
    ops.append(dict(kind="maxpool", src=x, dst=dst))
    x = dst
    for li in range(1, 5):
        for bi, blk in enumerate(getattr(net, f"layer{li}")):
            p = f"layer{li}.{bi}."
            if isinstance(blk, ResNestBottleneck):
                sa = blk.conv2
                o1 = conv_bn(x, p + "conv1", p + "bn1", blk.conv1)
                x2 = conv_bn(o1, p + "conv2.conv", p + "conv2.bn0", sa.conv)
                gap = buf()
                ops.append(dict(kind="splat_gap", src=x2, dst=gap))
                a1 = conv_bn(gap, p + "conv2.fc1", p + "conv2.bn1", sa.fc1)
                logits = buf()
                ops.append(dict(kind="conv", src=a1, dst=logits, key=prefix + p + "conv2.fc2", stride=1, pad=0, groups=1, need_dx=True))
                o2 = buf()
                ops.append(dict(kind="splat_mix", src=x2, logits=logits, dst=o2, gap=gap))
                if blk.avd_last is not None:
                    d2 = buf()
                    ops.append(dict(kind="avgpool3", src=o2, dst=d2))
                    o2 = d2
                res = x
                if blk.downsample is not None:
                    if isinstance(blk.downsample[0], nn.AvgPool2d):
                        r = buf()
                        ops.append(dict(kind="avgpool2", src=x, dst=r))
                        res = conv_bn(r, p + "downsample.1", p + "downsample.2", blk.downsample[1], relu=False)
                    else:
                        res = conv_bn(x, p + "downsample.1", p + "downsample.2", blk.downsample[1], relu=False)
                x = conv_bn(o2, p + "conv3", p + "bn3", blk.conv3, relu=True, res=res)
            elif isinstance(blk, BasicBlock):
                o1 = conv_bn(x, p + "conv1", p + "bn1", blk.conv1)
                res = x
                if blk.downsample is not None:
                    res = conv_bn(x, p + "downsample.0", p + "downsample.1", blk.downsample[0], relu=False)
                x = conv_bn(o1, p + "conv2", p + "bn2", blk.conv2, relu=True, res=res)
            else:
                raise L.ScouterError(f"cannot lower block type {type(blk).__name__}")
    return ops, nbuf[0], x


def backward_schedule(ops, feature_buffer: int):
    """Reverse-mode schedule of a ``lower_backbone_train`` program.  Gradient buffer ids equal forward buffer ids; an
    entry's ``acc`` says whether its gradient output must be ADDED to a buffer some later consumer already wrote (a
    block input feeds both the main path and the shortcut) or may overwrite it.  ``splat_gap``'s backward is the
    split-attention *apply* stage: it needs the gradient of the mixed output (``d_mix``) and of the pooled descriptor."""
    written = {feature_buffer}                    # the head's backward delivers d(features)
    sched = []
    mix_of_gap = {op["gap"]: op for op in ops if op["kind"] == "splat_mix"}
    for op in reversed(ops):
        e = dict(op)
        outs = []
        if op["kind"] == "conv":
            outs = [op["src"]] if op["need_dx"] else []
        elif op["kind"] == "bn":
            outs = [op["src"]] + ([op["res"]] if op["res"] >= 0 else [])
        elif op["kind"] == "splat_mix":
            outs = [op["logits"]]
        elif op["kind"] == "splat_gap":
            mix = mix_of_gap[op["dst"]]
            e.update(d_mix=mix["dst"], logits=mix["logits"])
            outs = [op["src"]]
        else:
            outs = [op["src"]]
        e["acc"] = {o: (o in written) for o in outs}
        written.update(outs)
        sched.append(e)
    return sched


class CompiledProgram:
    """A ``scouter_plan_t`` plus its bound arena for one input shape."""

    def __init__(self, program: Program, math: int):
        self.program = program
        arr = (L.Op * len(program.ops))(*program.ops)
        h = C.c_void_p()
        # plan_create reads the small stem filter banks back with a synchronous copy on the legacy stream: make sure the
        # fold kernels that produced them (torch's current stream, possibly a non-blocking one) have finished
        if torch.cuda.is_available() and any(t.is_cuda for t in program.keep):
            torch.cuda.current_stream().synchronize()
        L.check(L.lib().scouter_plan_create(arr, len(program.ops), program.nbuf, math, C.byref(h)), "scouter_plan_create")
        self.handle = h
        self.shape = None
        self.arena = None

    def bind(self, b, cin, h, w, device):
        if self.shape == (b, cin, h, w) and self.arena is not None and self.arena.device == device:
            return
        L.check(L.lib().scouter_plan_bind(self.handle, b, cin, h, w), "scouter_plan_bind")
        nbytes = L.lib().scouter_plan_arena_bytes(self.handle)
        self.arena = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        self.arena_off = (-self.arena.data_ptr()) % 1024
        self.arena_bytes = nbytes
        self.shape = (b, cin, h, w)

    @property
    def arena_ptr(self):
        return self.arena.data_ptr() + self.arena_off

    def buffer_shape(self, buf):
        s = (C.c_int32 * 4)()
        L.check(L.lib().scouter_plan_buffer_shape(self.handle, buf, C.byref(s)), "scouter_plan_buffer_shape")
        return tuple(s)

    def buffer_view(self, buf) -> torch.Tensor:
        """fp32 view (B,H,W,C) of a plan buffer inside the arena (valid until the next run overwrites it)."""
        b, h, w, c = self.buffer_shape(buf)
        off = L.lib().scouter_plan_buffer_offset(self.handle, buf)
        n = b * h * w * c
        return self.arena[self.arena_off + off: self.arena_off + off + 4 * n].view(torch.float32).view(b, h, w, c)

    def run(self, x: torch.Tensor):
        L.check(L.lib().scouter_plan_run(self.handle, x.data_ptr(), self.arena_ptr, self.arena_bytes, L.stream_ptr()),
                "scouter_plan_run")

    @property
    def launches(self):
        return L.lib().scouter_plan_launch_count(self.handle)

    def __del__(self):
        try:
            if self.handle:
                L.lib().scouter_plan_destroy(self.handle)
        except Exception:
            pass


def _check_inference(module: nn.Module, what: str):
    if module.training:
        raise NotImplementedError(
            f"{what}: train-mode forward (batch-statistics BatchNorm + backward) is SURVEY.md row f1 and is not "
            "implemented; call .eval() -- scouter_b200 accelerates the eval-mode forward hot path")


class BackboneRunner:
    """``ResNet.forward`` (resnet.py:503-509) on the CUDA library."""

    def __init__(self, net, math: int | None = None, fast_stages=None):
        self.net = net
        self.math = math
        self.fast_stages = fast_stages       # None: SCOUTER_TC_FAST_STAGES (default: none)
        self.sig = None
        self.cp = None

    def invalidate(self):
        """Drop the folded / packed weights (rebuilt on the next call).  Needed after in-place parameter writes that
        bypass the version counter (``.data``); ordinary assignments and in-place ops are detected automatically."""
        self.cp, self.sig = None, None

    def _compile(self):
        from .backbone import Identical
        from . import default_math
        net = self.net
        math = default_math() if self.math is None else self.math
        p, feat = lower_backbone(net, math, default_fast_stages() if self.fast_stages is None else self.fast_stages)
        self.flatten_nchw = isinstance(net.global_pool, Identical)
        if self.flatten_nchw:
            out = p.emit(L.OP_TO_NCHW, feat, p.buf())
        else:
            out = p.emit(L.OP_GAP, feat, p.buf())
            if not isinstance(net.fc, Identical):
                fc = net.fc
                w = fc.weight.detach().float().contiguous()
                b = fc.bias.detach().float().contiguous()
                out = p.emit(L.OP_CONV, out, p.buf(), cin=fc.in_features, cout=fc.out_features, k=1, w=w, b=b)
        self.out_buf = out
        self.cp = CompiledProgram(p, math)
        self.sig = _version_signature(net)

    def __call__(self, x):
        _require_cuda(x, "backbone")
        _check_inference(self.net, "backbone")
        if self.cp is None or self.sig != _version_signature(self.net):
            self._compile()
        x = x.contiguous()
        b, cin, h, w = x.shape
        with torch.cuda.device(x.device):
            self.cp.bind(b, cin, h, w, x.device)
            self.cp.run(x)
            out = self.cp.buffer_view(self.out_buf)
            # (B,H,W,C) view of a TO_NCHW buffer is really (B,C,H,W) memory; flatten(1) either way
            return out.reshape(b, -1).clone()
