"""Loader of the UNMODIFIED reference (wbw520/scouter ``sloter/`` + ``timm/``) from the git-ignored ``baseline/_ref/``
(filled by ``scripts/fetch_ref.sh``; falls back to ``/root/reference`` in the build container).

Used by ``bench.py --impl reference`` (the reference's own ``SlotModel.forward`` on the host cores) and by bench.py's
``gpu_eager_baseline`` leg (the same module ``.cuda()``, i.e. stock PyTorch eager / cuDNN / cuBLAS on the same B200 -- the
vendor bar of SURVEY.md 2.2).  It is a *baseline*, never part of the product path: nothing under ``scouter_b200/`` imports it.

Harness-side shims only, the reference files are not edited:
1. ``torch._six`` stub -- vendored timm does ``from torch._six import container_abcs`` (timm/models/layers/helpers.py:6);
2. ``torch.normal(mean, signed_std)`` -- torch 1.6 did not validate ``std`` (sloter/utils/slot_attention.py:20-25).
"""
from __future__ import annotations

import argparse
import collections.abc
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.environ.get("SCOUTER_REFERENCE_ROOT", ""), os.path.join(_HERE, "_ref"), "/root/reference"]


def reference_root() -> str | None:
    for c in CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "sloter", "slot_model.py")) and os.path.isdir(os.path.join(c, "timm")):
            return c
    return None


def install_shims(root: str):
    sys.dont_write_bytecode = True
    if "torch._six" not in sys.modules:
        six = types.ModuleType("torch._six")
        six.container_abcs = collections.abc
        six.string_classes = (str, bytes)
        six.int_classes = int
        sys.modules["torch._six"] = six
        torch._six = six
    if not getattr(torch.normal, "_scouter_shim", False):
        _orig = torch.normal

        def normal(mean, std, *a, **k):
            if torch.is_tensor(mean) and torch.is_tensor(std):
                return mean + std * torch.randn_like(std)
            return _orig(mean, std, *a, **k)

        normal._scouter_shim = True
        torch.normal = normal
    if root not in sys.path:
        sys.path.insert(0, root)


def reference_args(**over) -> argparse.Namespace:
    """The attributes ``sloter.slot_model.SlotModel(args)`` reads (train.py:18-79 after ``param_translation``)."""
    a = dict(model="resnest26d", dataset="ImageNet", channel=2048, num_classes=10, pre_trained=False,
             use_slot=True, use_pre=False, grad=False, loss_status=1, freeze_layers=0, hidden_dim=64,
             slots_per_class=1, power=2, to_k_layer=3, lambda_value=1.0, vis=False, vis_id=0, img_size=260)
    a.update(over)
    return argparse.Namespace(**a)


def load_reference_model(state_dict=None, feature_size=None, **args_over):
    """The reference's ``SlotModel`` in eval mode (``None`` if no reference tree is present).  ``feature_size`` overrides the
    hard-wired 9 (sloter/slot_model.py:61-64) for inputs other than 260x260."""
    root = reference_root()
    if root is None:
        return None
    install_shims(root)
    from sloter.slot_model import SlotModel  # noqa: the reference's own class
    m = SlotModel(reference_args(**args_over))
    if state_dict is not None:
        m.load_state_dict(state_dict)
    if feature_size is not None:
        m.feature_size = feature_size
    return m.eval()
