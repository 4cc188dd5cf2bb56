"""CPU: the oracle restatement reproduces the goldens dumped from the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle -- the reference has no tests of its own."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_err
from oracle import backbone as ob
from oracle import head as oh
from oracle.make_golden import head_inputs
from scouter_b200.synth import fill_state_dict, synth_images
import scouter_b200 as sb
from oracle.refshim import make_args


def test_pe_table_matches_reference():
    z = np.load("tests/golden/pe_sine.npz")
    for k in z.files:
        h, w = map(int, k[3:].split("x"))
        assert float(np.abs(oh.sine_pe(64, h, w).numpy() - z[k]).max()) < 1e-6


@pytest.mark.parametrize("name", golden_names("head"))
def test_head_oracle_vs_reference_golden(name):
    z, meta = load_golden(name)
    c = meta["case"]
    m = sb.SlotAttention(c["C"], c["spc"], 64, loss_status=c["ls"], power=c["power"], to_k_layer=c["L"])
    sd = fill_state_dict(m.state_dict(), seed=3)
    x_pe, x = head_inputs(c)
    logits, loss, attn = oh.xslot_forward(sd, x_pe, x, num_classes=c["C"], slots_per_class=c["spc"], loss_status=c["ls"],
                                          power=c["power"], return_attn=True)
    floor = rel_err(z["logits"], z["logits64"])
    assert rel_err(logits, z["logits"]) < max(1e-5, 20 * floor)
    assert abs(float(loss) - float(z["loss"])) < 1e-5
    afloor = float(np.abs(z["attn"] - z["attn64"]).max())
    assert float((attn - torch.from_numpy(z["attn"])).abs().max()) < max(1e-5, 20 * afloor)
    if z["vis"].size:
        v = oh.vis_maps_u8(torch.from_numpy(z["attn"]), num_classes=c["C"], slots_per_class=c["spc"])
        assert np.array_equal(v, z["vis"])


@pytest.mark.parametrize("name", golden_names("model"))
def test_model_oracle_vs_reference_golden(name):
    z, meta = load_golden(name)
    a = meta["args"]
    m = sb.SlotModel(make_args(**a))
    sd = fill_state_dict(m.state_dict(), seed=0)
    x = synth_images(meta["batch"], meta["cin"], meta["size"], meta["size"])
    tgt = torch.from_numpy(z["target"])
    o = ob.slot_model_forward(a["model"], sd, x, num_classes=a["num_classes"], slots_per_class=a["slots_per_class"],
                              loss_status=a["loss_status"], power=a["power"], lambda_value=a["lambda_value"], target=tgt,
                              return_attn=True)
    feat = ob.backbone_features(a["model"], sd, x)
    assert rel_err(feat[:, ::64], z["feat_sample"]) < 1e-5
    assert rel_err(o["logits"], z["logits"]) < 1e-4
    assert rel_err(o["log_probs"], z["log_probs"]) < 1e-4
    assert float((o["attn"] - torch.from_numpy(z["attn"])).abs().max()) < 1e-3
    got = np.array([float(o["loss"]), float(o["nll"]), float(o["attn_loss"])])
    assert np.allclose(got, z["losses"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("dataset", ["ImageNet", "MNIST"])
def test_preprocess_oracle_vs_reference_golden(dataset):
    """f2: the oracle's ToTensor+Normalize+cast equals what the reference's own transform objects produced, bit for bit."""
    z = np.load("tests/golden/preprocess_u8.npz")
    mean, std = sb.SlotModel.NORMALIZE[dataset]
    got = oh.preprocess_u8(z[dataset + "_u8"], mean, std).numpy()
    assert got.dtype == np.float32 and np.array_equal(got, z[dataset + "_f32"])


def test_vis_upsample_oracle_vs_pillow_golden():
    """f3: the numpy restatement of Pillow's 8-bit bilinear resampler equals Pillow's own output, bit for bit."""
    from conftest import vis_cases
    from oracle import vis as ov
    cases, meta = vis_cases()
    assert len(cases) == 8
    for name, (maps, heat, ratios) in cases.items():
        got = ov.resize_bilinear_u8(maps, heat.shape[1], heat.shape[2])
        assert got.dtype == np.uint8 and np.array_equal(got, heat), name
        assert np.array_equal(np.array([ov.attention_ratio(m) for m in maps]), ratios), name


def test_vis_upsample_oracle_vs_installed_pillow():
    """Same check against whatever Pillow is installed where the tests run, on fresh random maps and sizes."""
    Image = pytest.importorskip("PIL.Image")
    from oracle import vis as ov
    rng = np.random.RandomState(7)
    for h, w, oh_, ow in [(7, 7, 224, 224), (9, 9, 260, 260), (7, 9, 61, 35), (9, 9, 3, 4), (5, 5, 5, 40)]:
        m = rng.randint(0, 256, (h, w)).astype(np.uint8)
        ref = np.array(Image.fromarray(m).resize((ow, oh_), resample=Image.BILINEAR), dtype=np.uint8)
        assert np.array_equal(ov.resize_bilinear_u8(m, oh_, ow), ref), (h, w, oh_, ow)


@pytest.mark.parametrize("name", golden_names("train"))
def test_train_step_oracle_vs_reference_golden(name):
    """Row f1 bar (no CUDA backward exists yet): the train-mode restatement + autograd reproduces what the unmodified
    reference's ``model.train(); loss.backward()`` (engine.py:28-33) produced -- losses, every parameter's gradient
    (max-norm and L2 for all 273, element-wise for the head and a few backbone tensors) and the BatchNorm running
    statistics after the step.  Per-parameter bar: 1e-3 relative (floored at 1e-4 of the largest gradient), widened only
    by 4x the reference's own fp32-vs-fp64 floor on that parameter."""
    from oracle.train import train_step
    z, meta = load_golden(name)
    a = meta["args"]
    m = sb.SlotModel(make_args(**a))
    sd = fill_state_dict(m.state_dict(), seed=0)
    x = synth_images(meta["batch"], meta["cin"], meta["size"], meta["size"])
    tgt = torch.from_numpy(z["target"])
    o = train_step(a["model"], sd, x, tgt, num_classes=a["num_classes"], slots_per_class=a["slots_per_class"],
                   loss_status=a["loss_status"], power=a["power"], lambda_value=a["lambda_value"])
    assert rel_err(o["log_probs"], z["log_probs"]) < 1e-4
    got = np.array([float(o["loss"]), float(o["nll"]), float(o["attn_loss"])])
    assert np.allclose(got, z["losses"], rtol=1e-5, atol=1e-5)
    names = meta["param_names"]
    assert names == [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    scale = float(np.nanmax(z["grad_max"]))
    for i, n in enumerate(names):
        g = o["grads"][n]
        if np.isnan(z["grad_max"][i]):
            assert g is None and n.startswith("slot.to_q."), n          # unused parameter (train.py:140)
            continue
        bar = max(1e-3, 4 * float(z["grad_floor"][i]))
        den = max(float(z["grad_max"][i]), 1e-4 * scale)
        assert abs(float(g.abs().max()) - z["grad_max"][i]) / den < bar, n
        assert abs(float(g.double().norm()) - z["grad_l2"][i]) / max(z["grad_l2"][i], 1e-4 * scale) < bar, n
        if "grad." + n in z.files:
            ref = torch.from_numpy(z["grad." + n])
            gg = g[:, ::8] if n == "conv1x1.weight" else g
            assert float((gg - ref).abs().max()) / den < bar, n
    for k in [f for f in z.files if f.startswith("bn.")]:
        assert torch.equal(o["bn_updates"][k[3:]], torch.from_numpy(z[k])), k


@pytest.mark.parametrize("case", [(10, 1, 3, 1, 2, 3, 64, 7, 7), (5, 2, 1, -1, 1, 2, 32, 3, 4), (4, 3, 2, -1, 3, 1, 16, 2, 2)])
def test_hand_written_head_backward_equals_autograd(case):
    """oracle/head_backward.py (the explicit formulas a CUDA backward will implement) == autograd of oracle/head.py, fp64."""
    from oracle.head_backward import head_backward
    C, spc, L, ls, pw, B, ch, h, w = case
    m = sb.SlotModel(make_args(num_classes=C, slots_per_class=spc, to_k_layer=L, loss_status=ls, power=pw, channel=ch))
    sd = {k: v for k, v in fill_state_dict(m.state_dict(), seed=2).items() if not k.startswith("backbone.")}
    g = torch.Generator().manual_seed(11)
    params = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    feat = torch.randn(B, ch, h, w, dtype=torch.float64, generator=g).abs().requires_grad_(True)
    o = oh.head_forward(params, feat, num_classes=C, slots_per_class=spc, loss_status=ls, power=pw, dtype=torch.float64)
    gl, ga = torch.randn(B, C, dtype=torch.float64, generator=g), 0.7
    keys = [k for k in params if not k.startswith("slot.to_q")]
    grads = torch.autograd.grad((gl * o["logits"]).sum() + ga * o["attn_loss"], [feat] + [params[k] for k in keys])
    mine = head_backward(sd, feat, gl, ga, num_classes=C, slots_per_class=spc, loss_status=ls, power=pw)
    for k, ref in zip(["feat"] + keys, grads):
        assert rel_err(mine[k], ref) < 1e-10, k


def test_hand_written_head_backward_vs_reference_train_golden():
    """The same formulas, fed with the train-mode backbone features and d(loss)/d(logits) of nll(log_softmax), give the
    head gradients the unmodified reference's loss.backward() produced (train_cfg2 golden)."""
    from oracle.backbone import TrainState, backbone_features
    from oracle.head_backward import head_backward
    z, meta = load_golden("train_cfg2_resnest26d_224")
    a = meta["args"]
    m = sb.SlotModel(make_args(**a))
    sd = fill_state_dict(m.state_dict(), seed=0)
    x = synth_images(meta["batch"], meta["cin"], meta["size"], meta["size"])
    tgt = torch.from_numpy(z["target"])
    with torch.no_grad():
        feat = backbone_features(a["model"], TrainState(sd), x)
        o = oh.head_forward(sd, feat, num_classes=a["num_classes"], slots_per_class=a["slots_per_class"],
                            loss_status=a["loss_status"], power=a["power"], dtype=torch.float64)
    g_logits = (o["log_probs"].exp() - torch.nn.functional.one_hot(tgt, a["num_classes"])) / meta["batch"]   # d nll / d logits
    mine = head_backward(sd, feat, g_logits, a["lambda_value"], num_classes=a["num_classes"],
                         slots_per_class=a["slots_per_class"], loss_status=a["loss_status"], power=a["power"])
    checked = 0
    for k in [f[5:] for f in z.files if f.startswith("grad.") and not f.startswith("grad.backbone.")]:
        ref = torch.from_numpy(z["grad." + k])
        got = mine[k][:, ::8] if k == "conv1x1.weight" else mine[k]
        assert rel_err(got, ref) < 1e-3, k
        checked += 1
    assert checked == 13          # conv1x1 (2) + initial_slots + 3 to_k layers (6) + gru (4)
