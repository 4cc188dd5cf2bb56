"""CPU: the C-ABI library builds, loads and exports exactly what include/scouter_b200.h declares."""
import ctypes
import re

import pytest

from scouter_b200 import _lib as L


def declared_functions():
    src = open("include/scouter_b200.h").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scouter_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(L.SIGNATURES)


def test_library_loads_and_exports_every_symbol():
    lib = L.lib()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.scouter_abi_version() == 3


def test_argument_errors_without_gpu():
    lib = L.lib()
    assert lib.scouter_pe_sine(0, 64, 7, 7, 0) == -1
    assert b"pe_sine" in lib.scouter_last_error()
    d = L.XSlotDesc()
    d.d = 32
    assert lib.scouter_xslot_packed_bytes(ctypes.byref(d)) == 0            # hidden dim 32 unsupported
    assert b"hidden dim" in lib.scouter_last_error()
    h = ctypes.c_void_p()
    assert lib.scouter_plan_create(None, 0, 0, 0, ctypes.byref(h)) == -1


def test_head_launch_count_and_workspace_are_host_logic():
    """scouter_head_launch_count / _workspace_bytes validate their arguments and answer without touching a device:
    1 launch when the fused kernel applies (needs the driver's tensor-map encoder, i.e. a GPU box), else projection + loop."""
    lib = L.lib()
    d = L.XSlotDesc()
    d.d, d.num_classes, d.slots_per_class, d.to_k_layers, d.iters, d.loss_status, d.power = 64, 10, 1, 3, 3, 1, 2.0
    io = L.HeadIO()
    io.batch, io.h, io.w, io.channel, io.layout, io.math = 256, 7, 7, 2048, L.LAYOUT_NHWC, L.MATH_TC
    n = lib.scouter_head_launch_count(ctypes.byref(d), ctypes.byref(io))
    assert n in (1, 2, 3)                                  # 2 = fused kernel + on-the-fly weight split (no conv_w_split given)
    ws = lib.scouter_head_workspace_bytes(ctypes.byref(d), ctypes.byref(io))
    assert ws >= 4 * 256 * 49 * 64 * 4 + 2 * 64 * 2048 * 2  # split-K slabs + bf16 [W ; W_r] of conv1x1.weight
    io.channel = 2050                                      # not a multiple of 16
    assert lib.scouter_head_launch_count(ctypes.byref(d), ctypes.byref(io)) == 0
    assert lib.scouter_head_workspace_bytes(ctypes.byref(d), ctypes.byref(io)) == 0


def test_plan_shape_inference_on_cpu():
    """Shape inference and arena placement are host code: check both geometries without a GPU."""
    import scouter_b200 as sb
    from oracle.refshim import make_args
    from scouter_b200.plan import CompiledProgram, lower_backbone
    m = sb.SlotModel(make_args())
    prog, feat = lower_backbone(m.backbone)
    for p in prog.ops:   # weights are CPU tensors here; only pointers are recorded, nothing is launched
        pass
    cp = CompiledProgram(prog, L.MATH_FP32)
    for size, fs in ((224, 7), (260, 9)):
        L.check(L.lib().scouter_plan_bind(cp.handle, 4, 3, size, size))
        s = (ctypes.c_int32 * 4)()
        L.check(L.lib().scouter_plan_buffer_shape(cp.handle, feat, ctypes.byref(s)))
        assert tuple(s) == (4, fs, fs, 2048)
        assert L.lib().scouter_plan_arena_bytes(cp.handle) > 0
        assert L.lib().scouter_plan_launch_count(cp.handle) == len(prog.ops) + 8     # split-attention GAP = 2 launches


def test_plan_bind_refuses_sizes_that_overflow_32bit_indexing():
    """Unsupported sizes are a loud error of the C ABI (no silent wrap-around): >= 2^31 elements in any buffer, > 65535 images, wrong channel count."""
    import scouter_b200 as sb
    from oracle.refshim import make_args
    from scouter_b200.plan import CompiledProgram, lower_backbone
    prog, feat = lower_backbone(sb.SlotModel(make_args()).backbone)
    cp = CompiledProgram(prog, L.MATH_FP32)
    lib = L.lib()
    assert lib.scouter_plan_bind(cp.handle, 256, 3, 224, 224) == 0
    assert lib.scouter_plan_arena_bytes(cp.handle) < 4 << 30                   # ~2.6 GB at B=256
    assert lib.scouter_plan_bind(cp.handle, 4096, 3, 224, 224) == -2           # 4096 x 112 x 112 x 64 elements after conv1.6
    assert b"2^31" in lib.scouter_last_error()
    assert lib.scouter_plan_arena_bytes(cp.handle) == 0                        # a failed bind leaves the plan unbound
    assert lib.scouter_plan_bind(cp.handle, 70000, 3, 32, 32) == -2
    assert lib.scouter_plan_bind(cp.handle, 2, 4, 224, 224) == -1              # channel mismatch with the stem
    with pytest.raises(sb.ScouterError):
        L.check(lib.scouter_plan_bind(cp.handle, 0, 3, 224, 224))
    assert lib.scouter_plan_bind(cp.handle, 2, 3, 224, 224) == 0               # and the plan is still usable
