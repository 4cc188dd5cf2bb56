"""GPU (-m gpu): the convolution kernels one op at a time, through the C ABI (scouter_conv_forward),
against torch's CPU conv2d on the same seeded inputs (the oracle's primitive, oracle/backbone.py:_conv).

SCOUTER_MATH_TC (3xTF32, operands split on the fly) is held to the same 2e-5 as the exact CUDA-core kernel on
arbitrary fp32 operands; SCOUTER_MATH_TC_FAST (one tf32 pass) gets tf32-representable operands, as its producers
guarantee, and is checked to the tf32 rounding of its output.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from scouter_b200 import _lib as L
from scouter_b200.plan import round_tf32

pytestmark = pytest.mark.gpu

# (B, H, W, Cin, Cout, k, groups, residual, relu)   -- every distinct conv geometry of resnest26d at 224 and 260,
# ragged tiles, multi-image boxes, and the head projection
CASES = [
    (2, 56, 56, 64, 64, 1, 1, False, True),        # layer1.0.conv1
    (2, 56, 56, 64, 128, 3, 2, False, True),       # layer1.x.conv2.conv (grouped 3x3, N=64 per group)
    (3, 56, 56, 64, 256, 1, 1, True, True),        # layer1.x.conv3 + residual
    (2, 112, 112, 32, 32, 3, 1, False, True),      # stem conv1.3
    (1, 112, 112, 32, 64, 3, 1, False, True),      # stem conv1.6
    (2, 28, 28, 256, 512, 3, 2, False, True),      # layer3.0.conv2.conv
    (5, 14, 14, 512, 1024, 3, 2, False, True),     # layer4.0.conv2.conv
    (9, 7, 7, 512, 1024, 3, 2, False, True),       # layer4.1.conv2.conv: 2 images per 98-row box, odd batch
    (4, 7, 7, 2048, 512, 1, 1, False, True),       # layer4.1.conv1 (K = 2048)
    (4, 7, 7, 512, 2048, 1, 1, True, True),        # layer4.1.conv3 + residual (16 n-tiles)
    (2, 65, 65, 64, 128, 3, 2, False, True),       # 260^2 geometry: odd maps, ragged boxes
    (2, 33, 33, 128, 256, 3, 2, False, False),
    (3, 9, 9, 512, 1024, 3, 2, False, True),
    (1, 17, 17, 1024, 256, 1, 1, False, True),
    (1, 17, 23, 32, 32, 3, 1, True, False),        # folded-tap halo kernel: ragged 14 x 8 tiles, residual
    (2, 20, 20, 64, 64, 3, 1, False, True),        # folded-tap halo kernel: two channel blocks = two accumulation chunks
    (1, 5, 3, 32, 32, 3, 1, False, False),         # tiny map, smaller than one box
    (300, 1, 1, 64, 64, 1, 1, False, False),       # M = 300: partial last flat tile
    (1, 17, 17, 64, 128, 1, 1, True, True),        # residual epilogue with a partial last tile (M = 289)
    (2, 14, 14, 256, 320, 1, 1, True, False),      # residual, BN = 64, several k-blocks per chunk
]


def run_conv(dev, x, w, b, res, k, groups, relu, math, presplit=False):
    """x NCHW cpu, w (Cout,Cin/g,k,k) cpu -> out NCHW cpu via the library."""
    from scouter_b200.plan import split_weights_f16
    Bn, Cin, H, W = x.shape
    Cout = w.shape[0]
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
    wd = w.permute(0, 2, 3, 1).contiguous().to(dev)
    w2 = 0
    if presplit:                                      # [fp16 W ; bf16 (W - fp16 W)], like plan.py
        keep = split_weights_f16(wd)
        w2 = keep.data_ptr()
    bd = b.to(dev)
    rd = res.permute(0, 2, 3, 1).contiguous().to(dev) if res is not None else None
    out = torch.full((Bn, H, W, Cout), float("nan"), device=dev)
    op = L.Op(kind=L.OP_CONV, src=0, src2=-1, dst=1, cin=Cin, cout=Cout, kh=k, kw=k, stride=1, pad=k // 2, groups=groups,
              flags=(L.F_RELU if relu else 0) | (L.F_RESIDUAL if res is not None else 0), mid=0, reserved=0,
              w=wd.data_ptr(), b=bd.data_ptr(), w2=w2, b2=0)
    path = L.lib().scouter_conv_path(C.byref(op), Bn, H, W, math)
    L.check(L.lib().scouter_conv_forward(C.byref(op), xd.data_ptr(), L.ptr(rd), out.data_ptr(), Bn, H, W, math, 0))
    torch.cuda.synchronize()
    return out.permute(0, 3, 1, 2).cpu(), path


@pytest.mark.parametrize("case", CASES, ids=[f"B{c[0]}_{c[1]}x{c[2]}_c{c[3]}-{c[4]}_k{c[5]}g{c[6]}" for c in CASES])
@pytest.mark.parametrize("math", [L.MATH_FP32, L.MATH_TC, L.MATH_TC_FAST])
def test_conv_vs_torch_cpu(case, math):
    dev = torch.device("cuda", 0)
    Bn, H, W, Cin, Cout, k, groups, use_res, relu = case
    r = np.random.RandomState(hash(case) & 0xFFFF)
    x = torch.from_numpy(r.standard_normal((Bn, Cin, H, W)).astype(np.float32))
    w = torch.from_numpy((r.standard_normal((Cout, Cin // groups, k, k)) * np.sqrt(2.0 / (Cin // groups * k * k))).astype(np.float32))
    b = torch.from_numpy(0.1 * r.standard_normal(Cout).astype(np.float32))
    res = torch.from_numpy(r.standard_normal((Bn, Cout, H, W)).astype(np.float32)) if use_res else None
    if math == L.MATH_TC_FAST:
        x, w = round_tf32(x), round_tf32(w)          # what the producers of this mode hand to the kernel
    ref = F.conv2d(x.double(), w.double(), b.double(), 1, k // 2, 1, groups)
    if res is not None:
        ref = ref + res.double()
    if relu:
        ref = torch.relu(ref)
    out, path = run_conv(dev, x, w, b, res, k, groups, relu, math)
    assert path == (0 if math == L.MATH_FP32 else 1), "the tcgen05 kernel must be the one that runs in the TC modes"
    assert torch.isfinite(out).all()
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    if math == L.MATH_TC:
        # the plan's configuration: host-pre-split weights; 3x3 convs then take the halo kernel (path 2) unless the
        # map is too small to fill its 128 virtual rows
        out2, path2 = run_conv(dev, x, w, b, res, k, groups, relu, math, presplit=True)
        assert path2 in (1, 2) and (k == 1) == (path2 == 1) or min(H, W) <= 9
        err2 = float((out2.double() - ref).abs().max() / ref.abs().max())
        print(f"presplit path {path2}: err {err2:.2e} (on-the-fly split {err:.2e})")
        assert torch.isfinite(out2).all() and err2 < 5e-5, err2
    if math == L.MATH_TC_FAST:
        # the epilogue of this mode rounds its output to tf32 for the next layer
        assert err < 2.0 ** -11 * 1.5, err
        assert torch.all((out.view(torch.int32) & 0x1FFF) == 0)
    else:
        # exact fp32 FMA, and error-compensated 3xTF32 on arbitrary fp32 operands: both fp32-class (the TMEM
        # accumulator truncates instead of rounding, which costs ~n_mma * 2^-24 relative on long K)
        assert err < (2e-5 if math == L.MATH_FP32 else 5e-5), err


@pytest.mark.parametrize("scale", [1.0, 1e-2, 1e-4, 3e4])
def test_conv_tc_activation_scale(scale):
    """The activation remainder a - fp16(a) is carried in fp16 (csrc/ptx.cuh split2_act): exact to 2^-25 absolute, so the
    compensated product keeps its ~8e-7 for activations of ordinary scale and degrades gracefully (absolute error 2^-25 |w| per
    product) for tiny ones; values beyond fp16's range saturate the main operand and lose precision the same graceful way.
    Measured on B200 and asserted here: 8e-7 at scale 1, ~3e-6 at 1e-2, ~2e-4 at 1e-4 (still below one-pass tf32's 7e-4)."""
    dev = torch.device("cuda", 0)
    r = np.random.RandomState(5)
    x = torch.from_numpy((scale * r.standard_normal((2, 64, 28, 28))).astype(np.float32))
    w = torch.from_numpy((r.standard_normal((128, 64, 3, 3)) * np.sqrt(2.0 / 576)).astype(np.float32))
    b = torch.zeros(128)
    ref = F.conv2d(x.double(), w.double(), None, 1, 1)
    out, path = run_conv(dev, x, w, b, None, 3, 1, False, L.MATH_TC, presplit=True)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    print(f"activation scale {scale:g}: err {err:.2e}")
    assert err < {1.0: 2e-6, 1e-2: 1e-5, 1e-4: 6e-4, 3e4: 1e-3}[scale], err


def test_head_projection_is_fp32_accurate_on_tensor_cores():
    """conv1x1 + ReLU of the fused head (K = 2048) with the 3xTF32 kernel on raw fp32 features."""
    import scouter_b200 as sb
    dev = torch.device("cuda", 0)
    r = np.random.RandomState(11)
    Bn, n, ch = 37, 49, 2048
    feat = torch.from_numpy(np.maximum(r.standard_normal((Bn, n, ch)), 0).astype(np.float32))
    w = torch.from_numpy((r.standard_normal((64, ch)) / np.sqrt(ch)).astype(np.float32))
    b = torch.from_numpy(0.1 * r.standard_normal(64).astype(np.float32))
    ref = torch.relu(feat.double() @ w.double().t() + b.double())
    m = sb.SlotAttention(10, 1, 64, to_k_layer=3, power=2).to(dev).eval()
    desc, packed = m.desc_and_pack(dev)
    for math in (L.MATH_FP32, L.MATH_TC):
        io = L.HeadIO()
        io.batch, io.h, io.w, io.channel, io.layout, io.math = Bn, 7, 7, ch, L.LAYOUT_NHWC, math
        fd, wd, bd = feat.to(dev), w.to(dev), b.to(dev)
        pe = sb.build_position_encoding("sine", 64).table(7, 7, dev)
        logits = torch.empty(Bn, 10, device=dev)
        xo = torch.empty(Bn, n, 64, device=dev)
        io.feat, io.conv_w, io.conv_b, io.pe = fd.data_ptr(), wd.data_ptr(), bd.data_ptr(), pe.data_ptr()
        io.logits, io.attn, io.attn_sum, io.x_out = logits.data_ptr(), 0, 0, xo.data_ptr()
        nbytes = L.lib().scouter_head_workspace_bytes(C.byref(desc), C.byref(io))
        ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 1024
        L.check(L.lib().scouter_head_forward(C.byref(desc), packed.data_ptr(), C.byref(io), ws.data_ptr() + off, nbytes, 0))
        torch.cuda.synchronize()
        err = float((xo.cpu().double() - ref).abs().max() / ref.abs().max())
        print(f"head projection math={math}: rel err {err:.2e}")
        assert err < 2e-5
