"""scouter_b200 -- the SCOUTER (wbw520/scouter) forward hot path, rebuilt for B200 (sm_100a).

Public surface = the reference's module API for that path:
``SlotModel``, ``SlotAttention`` (alias ``ScouterAttention``), ``load_backbone``, ``Identical``,
``build_position_encoding`` / ``PositionEmbeddingSine`` and ``create_model`` for the hot-path backbones.
All arithmetic runs in ``libscouter_b200.so`` (hand-written CUDA behind the C ABI of
``include/scouter_b200.h``); there is no CPU or torch-op fallback.
"""
import os

from . import _lib
from ._lib import ScouterError, MATH_FP32, MATH_TC, MATH_TC_FAST
from .backbone import Identical, create_model, list_models
from .position_encode import PositionEmbeddingSine, build_position_encoding
from .slot_attention import ScouterAttention, SlotAttention
from .slot_model import SlotModel, load_backbone

__all__ = ["SlotModel", "SlotAttention", "ScouterAttention", "load_backbone", "Identical", "create_model",
           "list_models", "PositionEmbeddingSine", "build_position_encoding", "ScouterError", "default_math",
           "MATH_FP32", "MATH_TC", "MATH_TC_FAST"]


def default_math() -> int:
    """SCOUTER_MATH=tc (default): tcgen05 tensor cores, error-compensated products on fp32 data (fp32-class results);
    fp32: exact CUDA-core kernels; tc_fast: single-pass tf32 (cuDNN-TF32-class accuracy)."""
    v = os.environ.get("SCOUTER_MATH", "tc").lower()
    table = {"fp32": MATH_FP32, "tc": MATH_TC, "tc_fast": MATH_TC_FAST}
    if v not in table:
        raise ScouterError(f"SCOUTER_MATH={v!r}: expected one of {sorted(table)}")
    return table[v]
