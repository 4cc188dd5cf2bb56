"""Oracle (test infrastructure, see oracle/__init__.py): eval-mode backbones on the CPU.

Functional restatement, driven only by a reference-format ``state_dict``, of

* ``resnest26d`` trunk   -- reference ``timm/models/resnet.py:491-501`` (deep stem, max-pool, 4 stages),
                            ``timm/models/resnest.py:111-143`` (bottleneck),
                            ``timm/models/layers/split_attn.py:54-80`` (split attention, radix 2,
                            cardinality 1), ``timm/models/resnet.py:292-306`` (avg-down shortcut);
* ``resnet18`` trunk     -- reference ``timm/models/resnet.py:172-199`` (BasicBlock),
                            ``:276-289`` (conv shortcut), with the MNIST stem swap of
                            ``sloter/slot_model.py:23-24`` (1-channel 3x3 s2 conv instead of the 7x7).

Eval mode (BatchNorm uses running statistics, eps 1e-5) is what the shipped CUDA path is checked against.  For row f1
(backward + train-mode BatchNorm, not built yet) the same restatement runs in TRAIN mode when the
``state_dict`` is wrapped in ``TrainState``: BatchNorm then uses batch statistics (biased variance for the
normalisation, unbiased for the running update, momentum 0.1 -- ``nn.BatchNorm2d`` defaults, resnet.py:401-420) and the
updated running statistics are collected; everything is plain differentiable torch, so ``oracle/train.py`` gets the
reference's gradients from autograd.  The arithmetic primitives are
PyTorch's (``F.conv2d`` ...), exactly as in the reference, whose numeric kernel library (torch) is a
dependency that is not vendored under /root/reference (requirements.txt:27-28 pin torch==1.6.0).
Pinned against the imported reference by ``oracle/make_golden.py`` / ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


class TrainState(dict):
    """A ``state_dict`` whose BatchNorms run in training mode; ``updates`` collects the new running statistics."""

    def __init__(self, *a, momentum: float = 0.1, **k):
        super().__init__(*a, **k)
        self.momentum = momentum
        self.updates = {}


def _bn(sd, p, x, dtype):
    if isinstance(sd, TrainState):
        rm = sd[p + ".running_mean"].detach().to(dtype).clone()
        rv = sd[p + ".running_var"].detach().to(dtype).clone()
        y = F.batch_norm(x, rm, rv, sd[p + ".weight"].to(dtype), sd[p + ".bias"].to(dtype), True, sd.momentum, 1e-5)
        sd.updates[p + ".running_mean"], sd.updates[p + ".running_var"] = rm, rv
        return y
    return F.batch_norm(x, sd[p + ".running_mean"].to(dtype), sd[p + ".running_var"].to(dtype),
                        sd[p + ".weight"].to(dtype), sd[p + ".bias"].to(dtype), False, 0.0, 1e-5)


def _conv(sd, p, x, dtype, stride=1, padding=0, groups=1):
    bias = sd.get(p + ".bias")
    return F.conv2d(x, sd[p + ".weight"].to(dtype), None if bias is None else bias.to(dtype),
                    stride, padding, 1, groups)


def _split_attn(sd, p, x, dtype):
    """split_attn.py:54-80 with radix=2, groups(cardinality)=1."""
    x = torch.relu(_bn(sd, p + ".bn0", _conv(sd, p + ".conv", x, dtype, 1, 1, 2), dtype))
    b, rc, h, w = x.shape
    xr = x.reshape(b, 2, rc // 2, h, w)
    gap = xr.sum(1).mean((2, 3), keepdim=True)
    a = torch.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".fc1", gap, dtype), dtype))
    a = _conv(sd, p + ".fc2", a, dtype)                       # (B, 2C, 1, 1), radix-major
    a = torch.softmax(a.reshape(b, 2, rc // 2), dim=1)        # RadixSoftmax, cardinality 1
    return (xr * a[:, :, :, None, None]).sum(1)


def _resnest_block(sd, p, x, dtype, avd: bool, has_down: bool, down_pool: bool):
    out = torch.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x, dtype), dtype))
    out = _split_attn(sd, p + ".conv2", out, dtype)
    if avd:                                                   # avd_last, resnest.py:101
        out = F.avg_pool2d(out, 3, 2, 1)
    out = _bn(sd, p + ".bn3", _conv(sd, p + ".conv3", out, dtype), dtype)
    res = x
    if has_down:                                              # resnet.py:292-306
        if down_pool:
            res = F.avg_pool2d(res, 2, 2, ceil_mode=True, count_include_pad=False)
        res = _bn(sd, p + ".downsample.2", _conv(sd, p + ".downsample.1", res, dtype), dtype)
    return torch.relu(out + res)


def resnest26d_features(sd: dict, x: torch.Tensor, prefix: str = "backbone.", dtype=torch.float32, layers=(2, 2, 2, 2)):
    """(B,3,H,W) -> (B,2048,h,w).  ``layers=[2,2,2,2]`` (resnest.py:161-173); resnest14d / resnest50d differ only in
    the block counts ([1,1,1,1] :147-158, [3,4,6,3] :176-189)."""
    p = prefix
    x = x.to(dtype)
    x = torch.relu(_bn(sd, p + "conv1.1", _conv(sd, p + "conv1.0", x, dtype, 2, 1), dtype))
    x = torch.relu(_bn(sd, p + "conv1.4", _conv(sd, p + "conv1.3", x, dtype, 1, 1), dtype))
    x = torch.relu(_bn(sd, p + "bn1", _conv(sd, p + "conv1.6", x, dtype, 1, 1), dtype))
    x = F.max_pool2d(x, 3, 2, 1)
    for li in range(1, 5):
        for bi in range(layers[li - 1]):
            first = bi == 0
            x = _resnest_block(sd, f"{p}layer{li}.{bi}", x, dtype,
                               avd=first and li > 1, has_down=first, down_pool=first and li > 1)
    return x


def _basic_block(sd, p, x, dtype, stride: int, has_down: bool):
    out = torch.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x, dtype, stride, 1), dtype))
    out = _bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out, dtype, 1, 1), dtype)
    res = x
    if has_down:                                              # downsample_conv: 1x1 stride s, pad 0
        res = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x, dtype, stride, 0), dtype)
    return torch.relu(out + res)


def resnet18_features(sd: dict, x: torch.Tensor, prefix: str = "backbone.", dtype=torch.float32, layers=(2, 2, 2, 2)):
    """(B,Cin,H,W) -> (B,512,h,w).  The stem is whatever ``conv1.weight`` says: the MNIST swap
    (slot_model.py:23-24) is a 3x3 s2 p1 conv, the stock stem a 7x7 s2 p3 conv."""
    p = prefix
    x = x.to(dtype)
    k = sd[p + "conv1.weight"].shape[-1]
    x = torch.relu(_bn(sd, p + "bn1", _conv(sd, p + "conv1", x, dtype, 2, k // 2), dtype))
    x = F.max_pool2d(x, 3, 2, 1)
    for li in range(1, 5):
        for bi in range(layers[li - 1]):
            first = bi == 0 and li > 1
            x = _basic_block(sd, f"{p}layer{li}.{bi}", x, dtype, 2 if first else 1, first)
    return x


def backbone_features(model: str, sd: dict, x: torch.Tensor, dtype=torch.float32):
    resnest = {"resnest14d": (1, 1, 1, 1), "resnest26d": (2, 2, 2, 2), "resnest50d": (3, 4, 6, 3)}
    resnet = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3)}
    if model in resnest:
        return resnest26d_features(sd, x, dtype=dtype, layers=resnest[model])
    if model in resnet:
        return resnet18_features(sd, x, dtype=dtype, layers=resnet[model])
    raise ValueError(f"oracle has no restatement for backbone {model!r}")


def slot_model_forward(model: str, sd: dict, x: torch.Tensor, *, num_classes: int, slots_per_class: int,
                       loss_status: int = 1, power: int = 1, lambda_value: float = 1.0, target=None,
                       dtype=torch.float32, return_attn: bool = False):
    """sloter/slot_model.py:105-127 with ``feature_size`` taken from the actual feature map (D6)."""
    from .head import head_forward
    feat = backbone_features(model, sd, x, dtype)
    return head_forward(sd, feat, num_classes=num_classes, slots_per_class=slots_per_class,
                        loss_status=loss_status, power=power, lambda_value=lambda_value, target=target,
                        dtype=dtype, return_attn=return_attn)
