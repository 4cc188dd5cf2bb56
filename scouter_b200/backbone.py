"""Backbone zoo mirror: the module tree (names, shapes, ``state_dict`` keys) of the reference's vendored
timm models that the SCOUTER hot path uses, with the arithmetic done by libscouter_b200.

Reference counterparts:
  ``ResNet``              timm/models/resnet.py:380-509   (deep stem / plain stem, max-pool, 4 stages, pool + fc)
  ``ResNestBottleneck``   timm/models/resnest.py:58-143
  ``SplitAttnConv2d``     timm/models/layers/split_attn.py:31-80  (+ RadixSoftmax :14-28)
  ``BasicBlock``          timm/models/resnet.py:134-199
  ``create_model``        timm/models/factory.py:6-67 + registry.py:14-38 (only the hot-path names)

The classes are *parameter containers*: ``nn.Conv2d`` / ``nn.BatchNorm2d`` leaves hold the weights under
the reference's key names so checkpoints, ``dfs_freeze`` and DDP see the same tree; ``forward`` lowers the
tree to an op program (``scouter_b200.plan``) executed by the CUDA library.  There is no torch-op path.
"""
from __future__ import annotations

import math

import torch
from torch import nn


class Identical(nn.Module):
    """sloter/slot_model.py:10-15."""

    def forward(self, x):
        return x


class SplitAttnConv2d(nn.Module):
    """Split-attention conv, radix 2 / cardinality 1 as ResNeSt-d uses it (split_attn.py:31-80)."""

    def __init__(self, in_channels, out_channels, radix=2, reduction_factor=4):
        super().__init__()
        if radix != 2:
            raise NotImplementedError("scouter_b200 implements radix=2 split attention (resnest*d)")
        self.radix = radix
        mid_chs = out_channels * radix
        attn_chs = max(in_channels * radix // reduction_factor, 32)
        self.conv = nn.Conv2d(in_channels, mid_chs, 3, 1, 1, groups=radix, bias=False)
        self.bn0 = nn.BatchNorm2d(mid_chs)
        self.act0 = nn.ReLU(inplace=True)
        self.fc1 = nn.Conv2d(out_channels, attn_chs, 1)
        self.bn1 = nn.BatchNorm2d(attn_chs)
        self.act1 = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(attn_chs, mid_chs, 1)


class ResNestBottleneck(nn.Module):
    """resnest.py:58-143 with radix=2, cardinality=1, base_width=64, avd=True, avd_first=False."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        gw = planes
        self.avd_stride = stride if stride > 1 else 0      # avd and (stride > 1 or is_first); is_first never set
        self.conv1 = nn.Conv2d(inplanes, gw, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(gw)
        self.act1 = nn.ReLU(inplace=True)
        self.avd_first = None
        self.conv2 = SplitAttnConv2d(gw, gw)
        self.bn2 = None
        self.act2 = None
        self.avd_last = nn.AvgPool2d(3, self.avd_stride, padding=1) if self.avd_stride > 0 else None
        self.conv3 = nn.Conv2d(gw, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.act3 = nn.ReLU(inplace=True)
        self.downsample = downsample

    def zero_init_last_bn(self):
        nn.init.zeros_(self.bn3.weight)


class BasicBlock(nn.Module):
    """resnet.py:134-199 (no attention / anti-alias / drop layers: resnet18 creates none)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.act1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.act2 = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def zero_init_last_bn(self):
        nn.init.zeros_(self.bn2.weight)


def _downsample_avg(cin, cout, stride):      # resnet.py:292-306
    pool = nn.Identity() if stride == 1 else nn.AvgPool2d(2, stride, ceil_mode=True, count_include_pad=False)
    return nn.Sequential(pool, nn.Conv2d(cin, cout, 1, 1, 0, bias=False), nn.BatchNorm2d(cout))


def _downsample_conv(cin, cout, stride):     # resnet.py:276-289 with kernel_size 1
    return nn.Sequential(nn.Conv2d(cin, cout, 1, stride, 0, bias=False), nn.BatchNorm2d(cout))


class _GlobalAvgPool(nn.Module):
    """Stand-in for timm's SelectAdaptivePool2d('avg') (no parameters; lowered to SCOUTER_OP_GAP)."""

    def feat_mult(self):
        return 1


class ResNet(nn.Module):
    """resnet.py:380-509.  ``forward`` returns what the reference's does: the classifier output, or --
    once ``global_pool`` / ``fc`` have been replaced by ``Identical`` (slot_model.py:38-40) -- the
    feature map flattened from NCHW."""

    def __init__(self, block, layers, num_classes=1000, in_chans=3, stem_width=64, stem_type="", avg_down=False,
                 zero_init_last_bn=True):
        super().__init__()
        self.num_classes = num_classes
        self.block_type = block
        deep = "deep" in stem_type
        self.inplanes = stem_width * 2 if deep else 64
        if deep:
            self.conv1 = nn.Sequential(
                nn.Conv2d(in_chans, stem_width, 3, 2, 1, bias=False), nn.BatchNorm2d(stem_width), nn.ReLU(inplace=True),
                nn.Conv2d(stem_width, stem_width, 3, 1, 1, bias=False), nn.BatchNorm2d(stem_width), nn.ReLU(inplace=True),
                nn.Conv2d(stem_width, self.inplanes, 3, 1, 1, bias=False))
        else:
            self.conv1 = nn.Conv2d(in_chans, self.inplanes, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(self.inplanes)
        self.act1 = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.avg_down = avg_down
        for i, (planes, nblk, stride) in enumerate(zip((64, 128, 256, 512), layers, (1, 2, 2, 2)), 1):
            setattr(self, f"layer{i}", self._make_layer(block, planes, nblk, stride))
        self.global_pool = _GlobalAvgPool()
        self.num_features = 512 * block.expansion
        self.fc = nn.Linear(self.num_features, num_classes)
        for m in self.modules():                           # resnet.py:447-458
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0.0)
        if zero_init_last_bn:
            for m in self.modules():
                if hasattr(m, "zero_init_last_bn"):
                    m.zero_init_last_bn()
        self._runner = None

    def _make_layer(self, block, planes, blocks, stride):  # resnet.py:460-477
        down = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            mk = _downsample_avg if self.avg_down else _downsample_conv
            down = mk(self.inplanes, planes * block.expansion, stride)
        seq = [block(self.inplanes, planes, stride, down)]
        self.inplanes = planes * block.expansion
        seq += [block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*seq)

    def get_classifier(self):
        return self.fc

    def forward(self, x):
        from .plan import BackboneRunner
        if self._runner is None:
            object.__setattr__(self, "_runner", BackboneRunner(self))
        return self._runner(x)

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_runner"] = None
        return st


_REGISTRY = {}


def register_model(fn):
    _REGISTRY[fn.__name__] = fn
    return fn


@register_model
def resnet18(pretrained=False, num_classes=1000, in_chans=3, **kw):
    return ResNet(BasicBlock, [2, 2, 2, 2], num_classes, in_chans, **kw)


@register_model
def resnet34(pretrained=False, num_classes=1000, in_chans=3, **kw):
    return ResNet(BasicBlock, [3, 4, 6, 3], num_classes, in_chans, **kw)


def _resnest(layers, num_classes, in_chans, **kw):
    return ResNet(ResNestBottleneck, layers, num_classes, in_chans, stem_type="deep", stem_width=32, avg_down=True, **kw)


@register_model
def resnest14d(pretrained=False, num_classes=1000, in_chans=3, **kw):
    return _resnest([1, 1, 1, 1], num_classes, in_chans, **kw)


@register_model
def resnest26d(pretrained=False, num_classes=1000, in_chans=3, **kw):
    return _resnest([2, 2, 2, 2], num_classes, in_chans, **kw)


@register_model
def resnest50d(pretrained=False, num_classes=1000, in_chans=3, **kw):
    return _resnest([3, 4, 6, 3], num_classes, in_chans, **kw)


def list_models():
    return sorted(_REGISTRY)


def create_model(model_name, pretrained=False, num_classes=1000, in_chans=3, **kwargs):
    """timm/models/factory.py:6-67 for the hot-path backbones.  ``pretrained=True`` needs the network
    (helpers.py:75 downloads by URL) and is refused: load a checkpoint with ``load_state_dict`` instead."""
    if model_name not in _REGISTRY:
        raise RuntimeError(f"Unknown model ({model_name}); scouter_b200 implements {list_models()}")
    if pretrained:
        raise RuntimeError("pretrained=True would download weights; no network here -- pass pre_trained=False and "
                           "load a checkpoint (state_dict keys are reference-compatible)")
    for k in ("drop_connect_rate", "drop_path_rate", "drop_block_rate", "bn_tf", "bn_momentum", "bn_eps"):
        kwargs.pop(k, None)
    return _REGISTRY[model_name](pretrained=False, num_classes=num_classes, in_chans=in_chans, **kwargs)
