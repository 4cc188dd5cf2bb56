"""CPU: host-side mirror of the reference module API -- state_dict contract, constructor arguments,
error behaviour, op-program lowering.  No compute calls (there is no GPU here)."""
import json
import re

import pytest
import torch
from torch import nn

import scouter_b200 as sb
from oracle.refshim import make_args
from scouter_b200 import _lib as L
from scouter_b200.plan import fold_conv_bn, lower_backbone
from scouter_b200.synth import fill_state_dict

KEYS = json.load(open("tests/golden/state_dict_keys.json"))
CASES = {
    "cfg1_mnist_resnet18_260": dict(model="resnet18", dataset="MNIST", channel=512, num_classes=10, slots_per_class=1,
                                    power=1, to_k_layer=1),
    "cfg2_resnest26d_pos_224": dict(),
    "cfg5_cub200x2_224": dict(num_classes=200, slots_per_class=2),
    "no_slot_resnest26d": dict(use_slot=False),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_state_dict_keys_and_shapes_match_reference(name):
    m = sb.SlotModel(make_args(**CASES[name]))
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == KEYS[name]
    assert list(got) == list(KEYS[name])            # same order too


def test_param_counts():
    m = sb.SlotModel(make_args())
    assert sum(p.numel() for p in m.parameters()) == 15193824          # 15.194 M
    assert sum(p.numel() for n, p in m.named_parameters() if not n.startswith("backbone.")) == 173376
    m = sb.SlotModel(make_args(**CASES["cfg1_mnist_resnet18_260"]))
    assert sum(p.numel() for p in m.parameters()) == 11234432


def test_attributes_and_aliases():
    m = sb.SlotModel(make_args())
    for a in ("backbone", "conv1x1", "slot", "position_emb", "use_slot", "channel", "slots_per_class", "feature_size",
              "lambda_value"):
        assert hasattr(m, a)
    assert sb.ScouterAttention is sb.SlotAttention
    assert isinstance(m.backbone.global_pool, sb.Identical) and isinstance(m.backbone.fc, sb.Identical)
    assert m.slot.num_slots == 10 and m.slot.scale == 0.125
    # to_q is constructed but unused, like the reference (slot_attention.py:27-29, 52-53)
    assert "slot.to_q.0.weight" in m.state_dict()


def test_synthetic_state_dict_roundtrip_is_strict():
    m = sb.SlotModel(make_args())
    sd = fill_state_dict(m.state_dict(), seed=0)
    m.load_state_dict(sd, strict=True)
    sd2 = fill_state_dict(m.state_dict(), seed=0)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)               # deterministic in (name, shape, seed)


def test_dfs_freeze_matches_reference_semantics():
    m = sb.SlotModel(make_args())
    m.dfs_freeze(m.backbone, 2)                                       # layer4, layer3 stay trainable
    assert all(p.requires_grad for p in m.backbone.layer4.parameters())
    assert not any(p.requires_grad for p in m.backbone.layer1.parameters())
    assert not any(p.requires_grad for p in m.backbone.conv1.parameters())


def test_errors_are_loud_not_fallbacks():
    m = sb.SlotModel(make_args()).eval()
    with pytest.raises(sb.ScouterError):                              # CPU tensors: no CPU path
        m(torch.zeros(1, 3, 224, 224))
    with pytest.raises(RuntimeError):
        sb.create_model("densenet121")
    with pytest.raises(RuntimeError):
        sb.create_model("resnest26d", pretrained=True)
    with pytest.raises(ValueError):
        sb.build_position_encoding("nope", 64)


def test_bn_folding_matches_conv_bn():
    torch.manual_seed(0)
    conv = nn.Conv2d(8, 6, 3, padding=1, groups=2, bias=True)
    bn = nn.BatchNorm2d(6).eval()
    bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2); bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_()
    x = torch.randn(2, 8, 5, 5)
    w, b = fold_conv_bn(conv, bn)
    y = torch.nn.functional.conv2d(x, w.permute(0, 3, 1, 2), b, padding=1, groups=2)
    assert torch.allclose(y, bn(conv(x)), atol=1e-5)


def test_lowering_op_counts_resnest26d():
    m = sb.SlotModel(make_args())
    prog, feat = lower_backbone(m.backbone)
    assert lower_backbone(m.backbone, L.MATH_TC)[0].ops[1].w != 0
    kinds = [o.kind for o in prog.ops]
    # 47 convs in the backbone = 3 stem + 8*(conv1, conv2.conv, conv3) + 4 shortcuts + 16 fc1/fc2
    assert kinds.count(L.OP_STEM_CONV) == 1 and kinds.count(L.OP_CONV) == 2 + 24 + 4 + 16
    assert kinds.count(L.OP_SPLAT_GAP) == 8 and kinds.count(L.OP_SPLAT_APPLY) == 8
    assert kinds.count(L.OP_MAXPOOL) == 1 and kinds.count(L.OP_AVGPOOL) == 3
    assert feat == prog.ops[-1].dst


def test_round_tf32_is_round_to_nearest_ties_away():
    from scouter_b200.plan import round_tf32
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -11 + 2 ** -20, -1.0 - 2 ** -11, 3.14159265, -2.5e-7, 0.0])
    r = round_tf32(x)
    assert torch.all((r.view(torch.int32) & 0x1FFF) == 0)                     # 13 low mantissa bits cleared
    assert r[0] == 1.0 and r[1] == 1.0 + 2 ** -10 and r[3] == -1.0 - 2 ** -10   # ties away from zero
    assert float((r - x).abs().max() / x.abs().max()) <= 2 ** -11


def test_reference_checkpoint_flow_stage1_to_stage2(tmp_path, monkeypatch):
    """Row f4, host side: the reference's two-stage recipe on its own checkpoint format.  Stage 1 (``use_slot False``)
    saves ``{'model': state_dict, ...}`` (train.py:188-194) with ``backbone.``-prefixed keys incl. ``backbone.fc.*``;
    stage 2 (``use_slot True, use_pre True``) strips the prefix and loads it into the trunk before pool/fc become
    ``Identical`` (slot_model.py:26-40); test.py:117-120 loads a stage-2 checkpoint with strict keys."""
    monkeypatch.chdir(tmp_path)
    stage1 = sb.SlotModel(make_args(use_slot=False))
    sd1 = fill_state_dict(stage1.state_dict(), seed=5)
    assert "backbone.fc.weight" in sd1 and all(k.startswith("backbone.") for k in sd1)
    (tmp_path / "saved_model").mkdir()
    torch.save({"model": sd1, "epoch": 3}, tmp_path / "saved_model" / "ImageNet_no_slot_checkpoint.pth")

    stage2 = sb.SlotModel(make_args(use_slot=True, use_pre=True))
    sd2 = stage2.state_dict()
    assert not any(k.startswith("backbone.fc.") for k in sd2)                 # fc is Identical after the load
    trunk = [k for k in sd1 if not k.startswith("backbone.fc.")]
    assert trunk and all(torch.equal(sd1[k], sd2[k]) for k in trunk)         # every trunk tensor came from stage 1

    torch.save({"model": fill_state_dict(sd2, seed=6), "epoch": 9, "args": None}, tmp_path / "ImageNet_use_slot_checkpoint.pth")
    fresh = sb.SlotModel(make_args())
    ck = torch.load(tmp_path / "ImageNet_use_slot_checkpoint.pth", map_location="cpu", weights_only=False)
    fresh.load_state_dict(ck["model"])                                        # strict
    assert all(torch.equal(v, ck["model"][k]) for k, v in fresh.state_dict().items())


@pytest.mark.parametrize("cin,cout,k,pad,groups", [(8, 6, 3, 1, 2), (16, 32, 1, 0, 1), (4, 4, 3, 1, 1), (6, 12, 3, 0, 3)])
def test_dgrad_weights_turn_the_data_gradient_into_a_forward_conv(cin, cout, k, pad, groups):
    """f1 groundwork: dX of a stride-1 conv == a forward conv of dY with ``dgrad_weights`` (flipped taps, channels swapped
    per group, padding k-1-pad), i.e. the backward-data pass can run on the forward tcgen05 kernels."""
    from scouter_b200.plan import dgrad_weights
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, cin, 9, 7, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cout, cin // groups, k, k, generator=g, dtype=torch.float64)
    y = torch.nn.functional.conv2d(x, w, None, 1, pad, 1, groups)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (dx,) = torch.autograd.grad(y, x, dy)
    w_ohwi = w.permute(0, 2, 3, 1).contiguous()                       # the library's weight layout
    wd = dgrad_weights(w_ohwi, groups)                                # (Cin, kh, kw, Cout/g)
    assert wd.shape == (cin, k, k, cout // groups)
    got = torch.nn.functional.conv2d(dy, wd.permute(0, 3, 1, 2), None, 1, k - 1 - pad, 1, groups)
    assert got.shape == dx.shape and torch.allclose(got, dx, rtol=1e-12, atol=1e-12)


def test_invalidate_covers_data_writes():
    """in-place writes through ``.data`` do not bump ``_version`` -- the signature cannot see them, so an explicit
    ``invalidate()`` must drop every derived copy (program, per-shape states, packed head block, bf16 weight split)."""
    import torch
    import scouter_b200 as sb
    from scouter_b200.plan import _version_signature
    from scouter_b200.synth import make_args
    m = sb.SlotModel(make_args())
    s0 = _version_signature(m.backbone)
    with torch.no_grad():
        m.backbone.layer1[0].conv1.weight.data.mul_(2.0)
    assert _version_signature(m.backbone) == s0              # the documented blind spot
    with torch.no_grad():
        m.backbone.layer1[0].conv1.weight.mul_(2.0)
    assert _version_signature(m.backbone) != s0              # ordinary in-place ops are seen
    m._prog, m._sig, m._conv_w_split_key = ("stale",), 123, ("stale",)
    m._states["k"] = object()
    m.slot._packed, m.slot._packed_sig = object(), ("stale",)
    m.invalidate()
    assert m._prog is None and m._sig is None and not m._states and m._conv_w_split_key is None
    assert m.slot._packed is None and m.slot._packed_sig is None
    m._states["k"] = object()
    m.release_states()
    assert not m._states


def test_split_weights_f16_reconstructs_fp32():
    """The pre-split conv operand [fp16(W) ; bf16(W - fp16(W))] (plan.split_weights_f16, csrc/ptx.cuh split2_wgt): the two
    halves add back to W within bf16's rounding of a 2^-12 remainder, tiny values keep their remainder (bf16 has fp32's
    exponent range) and values beyond fp16's range saturate the first half while the second carries the excess."""
    from scouter_b200.plan import split_weights_f16
    w = torch.randn(8, 3, 3, 32) * 0.05
    w[0, 0, 0, 0], w[0, 0, 0, 1], w[0, 0, 0, 2] = 1e-7, 70000.0, -3.0e-5
    s = split_weights_f16(w)
    assert s.dtype == torch.int16 and s.shape == (16, 3, 3, 32)
    h = s[:8].view(torch.float16).float()
    r = s[8:].view(torch.bfloat16).float()
    assert float(h[0, 0, 0, 1]) == 65504.0 and abs(float(r[0, 0, 0, 1]) - (70000.0 - 65504.0)) < 32.0
    err = (h + r - w).abs()
    err[0, 0, 0, 1] = 0.0
    # |remainder| <= max(2^-12 |w|, 2^-25) (half an fp16 ulp; 2^-25 below fp16's normal range), stored with bf16's 2^-9 rounding
    bound = 2.0 ** -8 * torch.maximum(2.0 ** -11 * w.abs(), torch.full_like(w, 2.0 ** -24))
    assert bool((err <= bound).all())
    assert float((err / w.abs().clamp_min(1e-30))[w.abs() > 1e-3].max()) < 2.0 ** -19


def test_precision_policy_flags_only_the_listed_stages(monkeypatch):
    """SCOUTER_TC_FAST_STAGES / SlotModel.fast_stages (the measured-and-rejected per-stage policy): ops of the listed
    stages carry F_TF32_1PASS and no pre-split operand; every other conv keeps w2; other math modes ignore the policy."""
    from scouter_b200 import plan as P
    m = sb.SlotModel(make_args(model="resnest26d"))
    prog, _ = P.lower_backbone(m.backbone, L.MATH_TC, ("stem", "layer1"))
    flagged = [o for o in prog.ops if o.flags & L.F_TF32_1PASS]
    convs = [o for o in prog.ops if o.kind == L.OP_CONV]
    assert flagged and all(not o.w2 for o in flagged if o.kind == L.OP_CONV)
    assert any(o.w2 for o in convs) and all(bool(o.w2) != bool(o.flags & L.F_TF32_1PASS) for o in convs)
    # stem (3 convs + max-pool) + layer1 (2 blocks): the first op after them is layer2's conv1 and is not flagged
    n_flagged = len(flagged)
    prog2, _ = P.lower_backbone(m.backbone, L.MATH_TC, ("stem",))
    assert 0 < sum(1 for o in prog2.ops if o.flags & L.F_TF32_1PASS) < n_flagged
    prog3, _ = P.lower_backbone(m.backbone, L.MATH_FP32, ("stem", "layer1"))
    assert not any(o.flags & L.F_TF32_1PASS for o in prog3.ops)
    monkeypatch.setenv("SCOUTER_TC_FAST_STAGES", "layer4, stem")
    assert P.default_fast_stages() == ("layer4", "stem")
    monkeypatch.setenv("SCOUTER_TC_FAST_STAGES", "layer9")
    with pytest.raises(L.ScouterError):
        P.default_fast_stages()
