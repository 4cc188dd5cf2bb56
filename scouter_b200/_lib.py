"""ctypes binding of ``libscouter_b200.so`` (the C ABI declared in ``include/scouter_b200.h``).

There is no fallback of any kind: if the library is missing or fails to load, every compute entry
point raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCOUTER_B200_LIB: load another build of the same library (debug variants built by scripts/, e.g. -DSCOUTER_PROF)
LIB_PATH = os.environ.get("SCOUTER_B200_LIB") or os.path.join(_HERE, "libscouter_b200.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["api.cu", "conv_simt.cu", "aux_kernels.cu", "xslot.cu", "xslot_fast.cu", "head_fused.cu", "stem_ts.cu", "vis.cu", "umma_conv.cu", "umma_halo.cu",
           # row f1 (training step): C-ABI glue + the kernel bodies of csrc/draft/
           "train.cu", "draft/head_backward.cu", "draft/adamw.cu", "draft/bn_train.cu", "draft/pool_splat_bwd.cu", "draft/conv_wgrad.cu"]

OK = 0
LAYOUT_NHWC, LAYOUT_NCHW = 0, 1
MATH_FP32, MATH_TC, MATH_TC_FAST = 0, 1, 2
MAX_TO_K_LAYERS = 8

OP_STEM_CONV, OP_CONV, OP_MAXPOOL, OP_AVGPOOL = 1, 2, 3, 4
OP_SPLAT_GAP, OP_SPLAT_APPLY, OP_GAP, OP_TO_NCHW = 5, 7, 8, 9
F_RELU, F_RESIDUAL, F_CEIL_MODE, F_COUNT_INCLUDE_PAD, F_AVD_POOL, F_TF32_1PASS = 1, 2, 4, 8, 16, 32

_fp = C.c_void_p  # device pointers travel as integers


class XSlotDesc(C.Structure):
    _fields_ = [
        ("d", C.c_int32), ("num_classes", C.c_int32), ("slots_per_class", C.c_int32),
        ("to_k_layers", C.c_int32), ("iters", C.c_int32), ("loss_status", C.c_int32), ("power", C.c_float),
        ("initial_slots", _fp),
        ("to_k_w", _fp * MAX_TO_K_LAYERS), ("to_k_b", _fp * MAX_TO_K_LAYERS),
        ("gru_w_ih", _fp), ("gru_w_hh", _fp), ("gru_b_ih", _fp), ("gru_b_hh", _fp),
    ]


class XSlotIO(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n", C.c_int32),
        ("x", _fp), ("x_sb", C.c_int64), ("x_sn", C.c_int64), ("x_sd", C.c_int64),
        ("x_pe", _fp), ("xpe_sb", C.c_int64), ("xpe_sn", C.c_int64), ("xpe_sd", C.c_int64),
        ("pe", _fp),
        ("logits", _fp), ("attn", _fp), ("attn_sum", _fp),
    ]


class HeadIO(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("channel", C.c_int32),
        ("layout", C.c_int32), ("math", C.c_int32),
        ("feat", _fp), ("conv_w", _fp), ("conv_b", _fp), ("pe", _fp),
        ("logits", _fp), ("attn", _fp), ("attn_sum", _fp), ("x_out", _fp), ("conv_w_split", _fp),
    ]


class Op(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("src", C.c_int32), ("src2", C.c_int32), ("dst", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("groups", C.c_int32),
        ("flags", C.c_int32), ("mid", C.c_int32), ("reserved", C.c_int32),
        ("w", _fp), ("b", _fp), ("w2", _fp), ("b2", _fp),
    ]


class ForwardHostArgs(C.Structure):
    _fields_ = [
        ("plan", C.c_void_p), ("desc", C.POINTER(XSlotDesc)), ("packed", _fp),
        ("head", HeadIO), ("feat_buffer", C.c_int32), ("reserved", C.c_int32),
        ("input_host", _fp), ("input_dev", _fp), ("input_bytes", C.c_size_t),
        ("arena", _fp), ("arena_bytes", C.c_size_t),
        ("head_workspace", _fp), ("head_workspace_bytes", C.c_size_t),
        ("target_dev", _fp), ("lambda_value", C.c_float), ("reserved2", C.c_float),
        ("log_probs_dev", _fp), ("losses_dev", _fp), ("log_probs_host", _fp), ("losses_host", _fp),
        ("stream", _fp),
    ]


# name -> (restype, argtypes); also the list tests/test_abi.py checks against include/scouter_b200.h
SIGNATURES = {
    "scouter_abi_version": (C.c_int, []),
    "scouter_last_error": (C.c_char_p, []),
    "scouter_device_check": (C.c_int, [C.c_int]),
    "scouter_pe_sine": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, _fp]),
    "scouter_xslot_packed_bytes": (C.c_size_t, [C.POINTER(XSlotDesc)]),
    "scouter_xslot_pack": (C.c_int, [C.POINTER(XSlotDesc), _fp, _fp]),
    "scouter_xslot_workspace_bytes": (C.c_size_t, [C.POINTER(XSlotDesc), C.c_int, C.c_int]),
    "scouter_xslot_forward": (C.c_int, [C.POINTER(XSlotDesc), _fp, C.POINTER(XSlotIO), _fp, C.c_size_t, _fp]),
    "scouter_head_finalize": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                        _fp, _fp, _fp]),
    "scouter_vis_maps_u8": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp]),
    "scouter_vis_upsample_u8": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp]),
    "scouter_conv_forward": (C.c_int, [C.POINTER(Op), _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "scouter_conv_path": (C.c_int, [C.POINTER(Op), C.c_int, C.c_int, C.c_int, C.c_int]),
    "scouter_head_workspace_bytes": (C.c_size_t, [C.POINTER(XSlotDesc), C.POINTER(HeadIO)]),
    "scouter_head_launch_count": (C.c_int, [C.POINTER(XSlotDesc), C.POINTER(HeadIO)]),
    "scouter_head_forward": (C.c_int, [C.POINTER(XSlotDesc), _fp, C.POINTER(HeadIO), _fp, C.c_size_t, _fp]),
    "scouter_plan_create": (C.c_int, [C.POINTER(Op), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "scouter_plan_destroy": (None, [C.c_void_p]),
    "scouter_plan_bind": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "scouter_plan_arena_bytes": (C.c_size_t, [C.c_void_p]),
    "scouter_plan_buffer_shape": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32 * 4)]),
    "scouter_plan_buffer_offset": (C.c_size_t, [C.c_void_p, C.c_int]),
    "scouter_plan_run": (C.c_int, [C.c_void_p, _fp, _fp, C.c_size_t, _fp]),
    "scouter_plan_launch_count": (C.c_int, [C.c_void_p]),
    "scouter_preprocess_u8": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), _fp, _fp]),
    "scouter_forward_host": (C.c_int, [C.POINTER(ForwardHostArgs)]),
    "scouter_stem_conv_forward": (C.c_int, [C.POINTER(Op), _fp, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    # row f1
    "scouter_train_bn_forward": (C.c_int, [C.c_void_p, _fp]),
    "scouter_train_bn_backward": (C.c_int, [C.c_void_p, _fp]),
    "scouter_train_conv_wgrad": (C.c_int, [C.c_void_p, _fp]),
    "scouter_train_conv_dgrad": (C.c_int, [C.c_void_p, _fp]),
    "scouter_train_pool_backward": (C.c_int, [C.c_void_p, C.c_int, _fp]),
    "scouter_pool_forward": (C.c_int, [C.c_int, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "scouter_train_splat_backward": (C.c_int, [C.c_void_p, C.c_int, _fp]),
    "scouter_splat_gap_scratch_floats": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "scouter_splat_gap_forward": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    "scouter_splat_apply_forward": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "scouter_train_head_backward_scratch_floats": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "scouter_train_head_backward": (C.c_int, [C.c_void_p, _fp]),
    "scouter_train_adamw_step": (C.c_int, [C.c_void_p, _fp]),
    "scouter_split_weights_f16": (C.c_int, [_fp, _fp, C.c_size_t, _fp]),
}


# ---- row f1: argument blocks of the training entries (mirror include/scouter_b200.h field by field) -----------------
_f32p, _f64p = C.c_void_p, C.c_void_p


class BnTrainArgs(C.Structure):
    _fields_ = [("M", C.c_longlong), ("C", C.c_int), ("x", _fp), ("sums", _fp), ("gamma", _fp), ("beta", _fp),
                ("running_mean", _fp), ("running_var", _fp), ("scale", _fp), ("shift", _fp), ("save_mean", _fp), ("save_rstd", _fp),
                ("eps", C.c_float), ("momentum", C.c_float), ("residual", _fp), ("y", _fp), ("relu", C.c_int)]


class BnBwdArgs(C.Structure):
    _fields_ = [("M", C.c_longlong), ("C", C.c_int), ("x", _fp), ("out", _fp), ("d_out", _fp), ("gamma", _fp), ("save_mean", _fp),
                ("save_rstd", _fp), ("sums", _fp), ("d_gamma", _fp), ("d_beta", _fp), ("coef", _fp), ("dx", _fp), ("d_residual", _fp),
                ("relu", C.c_int)]


class WgradArgs(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("B", "H", "W", "Cin", "Ho", "Wo", "Cout", "k", "stride", "pad", "groups")] + \
               [("x", _fp), ("dy", _fp), ("dw", _fp), ("db", _fp)]


class DgradArgs(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("B", "H", "W", "Cin", "Ho", "Wo", "Cout", "k", "stride", "pad", "groups")] + \
               [("dy", _fp), ("w", _fp), ("dx", _fp)]


class PoolBwdArgs(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("B", "H", "W", "C", "Ho", "Wo")] + [("x", _fp), ("dy", _fp), ("dx", _fp)]


class SplatBwdArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("HW", C.c_int), ("C", C.c_int), ("x2", _fp), ("d_out", _fp), ("att", _fp), ("d_att", _fp),
                ("d_logit", _fp), ("d_gap", _fp), ("d_x2", _fp)]


class HeadBwdArgs(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("B", "n", "ch", "S", "C", "spc", "L", "iters", "loss_status")] + \
               [("feat", _fp), ("conv_w", _fp), ("conv_b", _fp), ("pe", _fp),
                ("to_k_w", _fp * MAX_TO_K_LAYERS), ("to_k_b", _fp * MAX_TO_K_LAYERS),
                ("w_ih", _fp), ("w_hh", _fp), ("b_ih", _fp), ("b_hh", _fp), ("slots0", _fp), ("g_logits", _fp), ("attn_coef", _fp),
                ("d_feat", _fp), ("d_pre", _fp), ("g_conv_w", _fp), ("g_conv_b", _fp),
                ("g_to_k_w", _fp * MAX_TO_K_LAYERS), ("g_to_k_b", _fp * MAX_TO_K_LAYERS),
                ("g_w_ih", _fp), ("g_w_hh", _fp), ("g_b_ih", _fp), ("g_b_hh", _fp), ("g_slots0", _fp),
                ("scratch", _fp), ("scratch_per_image", C.c_size_t)]


class AdamWArgs(C.Structure):
    _fields_ = [("decay", C.c_float), ("one_minus_beta1", C.c_float), ("beta2", C.c_float), ("one_minus_beta2", C.c_float),
                ("eps", C.c_float), ("step_size", C.c_float), ("bias_correction2_sqrt", C.c_float), ("n", C.c_size_t),
                ("p", _fp), ("g", _fp), ("m", _fp), ("v", _fp)]


_lib = None


class ScouterError(RuntimeError):
    pass


def nvcc_command(out_path: str = LIB_PATH, defines: tuple = ()) -> list[str]:
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(os.path.dirname(_HERE), "include"), *defines,
            *[os.path.join(CSRC, s) for s in SOURCES], "-o", out_path]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA extension in-tree for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "scouter_b200.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = nvcc_command()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        print(" ".join(cmd))
        print(r.stdout, r.stderr)
    if r.returncode:
        raise ScouterError(f"nvcc failed ({r.returncode}):\n{r.stderr[-4000:]}")
    return LIB_PATH


def lib():
    """The loaded library.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ScouterError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). scouter_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().scouter_last_error().decode(errors="replace")
        kind = {-1: "invalid argument", -2: "unsupported", -3: "bad state"}.get(rc, f"CUDA error {rc}")
        raise ScouterError(f"{what or 'libscouter_b200'}: {kind}: {msg}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (0 for None)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
