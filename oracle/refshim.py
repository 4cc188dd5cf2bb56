"""Oracle (test infrastructure): import the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
``oracle/make_golden.py`` to pin the restatements in ``oracle/head.py`` / ``oracle/backbone.py``
and to produce ``tests/golden/*.npz``.  Harness-side shims only -- the reference tree is never
edited:

1. ``torch._six`` stub: vendored timm does ``from torch._six import container_abcs``
   (reference ``timm/models/layers/helpers.py:6``).
2. ``torch.normal(mean, signed_std)``: torch 1.6 did not validate ``std``; torch >= 2 raises
   (reference ``sloter/utils/slot_attention.py:20-25``).  Emulated as ``mean + std * N(0,1)``.
3. ``args`` built by hand (``train.get_args_parser`` needs uninstalled packages).
"""
from __future__ import annotations

import collections.abc
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("SCOUTER_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sloter", "slot_model.py"))


def install_shims():
    sys.dont_write_bytecode = True  # /root/reference is read-only
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
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


from scouter_b200.synth import make_args  # noqa: E402,F401  (the args Namespace builder lives with the synthetic inputs)


def reference_slot_model(**over):
    install_shims()
    from sloter.slot_model import SlotModel  # noqa: the reference's
    return SlotModel(make_args(**over)).eval()


def reference_slot_attention(*a, **k):
    install_shims()
    from sloter.utils.slot_attention import SlotAttention  # noqa: the reference's
    return SlotAttention(*a, **k).eval()


class capture_sigmoid:
    """Record every ``torch.sigmoid`` output during a reference call (the last (B,S,n) one is the
    final attention map, ``slot_attention.py:57``, which the reference does not return)."""

    def __enter__(self):
        self.outs = []
        self._orig = torch.sigmoid

        def rec(x):
            y = self._orig(x)
            self.outs.append(y.detach().clone())
            return y

        torch.sigmoid = rec
        return self

    def __exit__(self, *exc):
        torch.sigmoid = self._orig
        return False
