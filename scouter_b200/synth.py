"""Deterministic synthetic weights and inputs (no network: no datasets, no checkpoints).

``fill_state_dict`` overwrites every entry of a reference-format ``state_dict``
with values that depend only on (key name, shape, seed) through ``numpy.random.RandomState`` -- the
legacy generator whose streams numpy guarantees never to change -- so the build container (where
the goldens are generated from the reference) and the GPU box (where they are checked) see
bit-identical weights without shipping a 60 MB checkpoint.

BatchNorm affine and running statistics are randomised on purpose: the reference constructor zeroes
the last BN of every residual block (``timm/models/resnet.py:455-458``, ``resnest.py:108-109``), which
would make every conv2 / split-attention / conv3 weight invisible to a parity test.
"""
from __future__ import annotations

import argparse
import zlib

import numpy as np
import torch


def make_args(**over) -> argparse.Namespace:
    """The attributes ``SlotModel(args)`` reads, with the defaults of the reference's parser after
    ``param_translation`` (train.py:18-79) for the resnest26d + xSlot recipe (README.md:39-42); override by keyword."""
    a = dict(model="resnest26d", dataset="ImageNet", channel=2048, num_classes=10, pre_trained=False,
             use_slot=True, use_pre=False, grad=False, loss_status=1, freeze_layers=0, hidden_dim=64,
             slots_per_class=1, power=2, to_k_layer=3, lambda_value=1.0, vis=False, vis_id=0, img_size=260)
    a.update(over)
    return argparse.Namespace(**a)


def _rng(name: str, seed: int) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def _is_block_final_bn(name: str) -> bool:
    # bn3 of a ResNeSt/ResNet bottleneck, bn2 of a BasicBlock: keep the residual stream tame.
    parts = name.split(".")
    return len(parts) >= 3 and parts[-2] in ("bn3", "bn2") and "layer" in name and "conv2" not in name


def synth_tensor(name: str, shape, seed: int = 0) -> np.ndarray:
    r = _rng(name, seed)
    shape = tuple(shape)
    leaf = name.split(".")[-1]
    if leaf == "num_batches_tracked":
        return np.zeros(shape, dtype=np.int64)
    if leaf == "running_var":
        return r.uniform(0.5, 1.5, shape).astype(np.float32)
    if leaf == "running_mean":
        return (0.1 * r.standard_normal(shape)).astype(np.float32)
    if leaf == "initial_slots":
        mu = r.standard_normal((1, 1, shape[-1]))
        sg = r.standard_normal((1, 1, shape[-1]))
        return (mu + sg * r.standard_normal(shape)).astype(np.float32)
    if "gru" in name or "to_k" in name or "to_q" in name:
        bound = 1.0 / np.sqrt(shape[-1] if len(shape) > 1 else 64)
        return r.uniform(-bound, bound, shape).astype(np.float32)
    if len(shape) == 4:                                   # conv weight, He / fan-in
        fan_in = shape[1] * shape[2] * shape[3]
        return (np.sqrt(2.0 / fan_in) * r.standard_normal(shape)).astype(np.float32)
    if len(shape) == 2:                                   # classifier fc
        return (np.sqrt(1.0 / shape[1]) * r.standard_normal(shape)).astype(np.float32)
    if len(shape) == 1:
        if leaf == "weight":                              # every 1-D weight is a BatchNorm gamma
            lo, hi = (0.2, 0.5) if _is_block_final_bn(name) else (0.5, 1.5)
            return r.uniform(lo, hi, shape).astype(np.float32)
        return (0.1 * r.standard_normal(shape)).astype(np.float32)   # biases
    return r.standard_normal(shape).astype(np.float32)


def fill_state_dict(sd: dict, seed: int = 0) -> dict:
    """New dict with the same keys/shapes/dtypes as ``sd`` and deterministic synthetic values."""
    out = {}
    for k, v in sd.items():
        t = torch.from_numpy(synth_tensor(k, v.shape, seed))
        out[k] = t.to(v.dtype) if v.dtype != torch.int64 else t.reshape(v.shape)
    return out


def synth_images(b: int, c: int, h: int, w: int, seed: int = 1234) -> torch.Tensor:
    """(B,C,H,W) fp32 ~N(0,1): a normalised image batch (reference ``dataset/transform_func.py:102-110``)."""
    r = np.random.RandomState(seed)
    return torch.from_numpy(r.standard_normal((b, c, h, w)).astype(np.float32))


def synth_labels(b: int, num_classes: int, seed: int = 1234) -> torch.Tensor:
    return torch.from_numpy(np.random.RandomState(seed + 1).randint(0, num_classes, size=(b,)).astype(np.int64))
