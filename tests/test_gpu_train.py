"""GPU (-m gpu): row f1, the training step through the C ABI (scouter_b200/train.py -> scouter_train_* entries) against
the vectors the UNMODIFIED reference produced with ``model.train(); loss.backward()`` (engine.py:28-33;
tests/golden/train_*.npz, oracle/make_golden_train.py).

What the bar can be.  The goldens record, per parameter, the reference's OWN fp32-vs-fp64 gradient noise ("floor"): the
eps-free sum-normalisation of slot_attention.py:56 amplifies rounding noise in the backward exactly as in the forward
(SURVEY D9), and ReLU masks / max-pool arg-maxes flip on near-ties.  With random weights the resnest26d floors are 1e-2
(median 3e-3); resnet18/MNIST is well conditioned (floors ~5e-5).  A second fp32 implementation of the same math is another
draw from that noise -- measured on B200 (this file prints the distribution): resnest26d errors sit at 1-2.6x the floor, the
well-conditioned resnet18 at 1.2e-4 median / 1.9e-3 worst of the largest gradient.  Hence, relative to max(|g_ref|max, 1e-4 of
the largest gradient), on max-norm, L2 norm and (where stored) every element:
    every parameter       err < max(1e-2,   8 x floor)   (ONE parameter per model may reach twice that, see below)
    90 % of parameters    err < max(2.5e-3, 4 x floor)
The one-parameter allowance: the split-attention fc1 / fc2 gradients pass through a BatchNorm whose batch statistics are taken
over just B = 4 values per channel (split_attn.py:66-67 on a (B,C,1,1) map), so they are almost pure amplified rounding noise
(floors 2e-3 .. 2e-2 of the gradient's own maximum) and a single draw lands at ~15 x floor now and then -- measured:
layer1.0.conv2.fc1.weight of cfg 3 at 3.1e-2 against a floor of 2.1e-3 with the fp16-main-product convs (inside its bar with round 1's
tf32 main product: a different draw of the same noise).
Log-probs and losses are held to the forward's bar.  The conv gradients are merged with fp32 atomics (order not fixed).
"""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from oracle.refshim import make_args
import scouter_b200 as sb
from scouter_b200 import _lib as L
from scouter_b200.synth import fill_state_dict, synth_images

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    L.check(L.lib().scouter_device_check(0))
    return torch.device("cuda", 0)


def build(meta, dev, math):
    m = sb.SlotModel(make_args(**meta["args"]))
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=0))
    m = m.to(dev).train()
    m.math = math
    return m


@pytest.mark.parametrize("math", [L.MATH_FP32, L.MATH_TC])
@pytest.mark.parametrize("name", golden_names("train"))
def test_train_step_vs_reference_golden(dev, name, math):
    z, meta = load_golden(name)
    m = build(meta, dev, math)
    x = synth_images(meta["batch"], meta["cin"], meta["size"], meta["size"]).to(dev)
    tgt = torch.from_numpy(z["target"]).to(dev)
    out, (loss, nll, attn_loss) = m(x, tgt)
    assert loss.requires_grad and not out.requires_grad
    loss.backward()
    torch.cuda.synchronize()
    tol = 2e-5 if math == L.MATH_FP32 else 1e-3
    lp_err = float((out.cpu() - torch.from_numpy(z["log_probs"])).abs().max() / max(1.0, float(np.abs(z["log_probs"]).max())))
    got = np.array([float(loss), float(nll), float(attn_loss)])
    assert lp_err < max(tol, 1e-4), lp_err
    assert np.allclose(got, z["losses"], rtol=10 * tol, atol=10 * tol), (got, z["losses"])
    names = meta["param_names"]
    params = dict(m.named_parameters())
    assert names == list(params)
    scale = float(np.nanmax(z["grad_max"]))
    worst = ("", 0.0)
    ratios, tight, over = [], 0, []
    for i, n in enumerate(names):
        g = params[n].grad
        if np.isnan(z["grad_max"][i]):
            assert g is None and n.startswith("slot.to_q."), n          # unused parameter (train.py:140 find_unused_parameters)
            continue
        assert g is not None, n
        g = g.cpu()
        bar = max(1e-2, 8 * float(z["grad_floor"][i]))
        den = max(float(z["grad_max"][i]), 1e-4 * scale)
        e1 = abs(float(g.abs().max()) - z["grad_max"][i]) / den
        e2 = abs(float(g.double().norm()) - z["grad_l2"][i]) / max(z["grad_l2"][i], 1e-4 * scale)
        e3 = 0.0
        if "grad." + n in z.files:
            ref = torch.from_numpy(z["grad." + n])
            gg = g[:, ::8] if n == "conv1x1.weight" else g
            e3 = float((gg - ref).abs().max()) / den
        e = max(e1, e2, e3)
        w = e / bar
        ratios.append((e / max(float(z["grad_floor"][i]), 1e-12), e, float(z["grad_floor"][i]), n))
        tight += e < max(2.5e-3, 4 * float(z["grad_floor"][i]))
        if w > worst[1]:
            worst = (n, w)
        assert e < 2 * bar, (n, e1, e2, e3, bar)
        if not e < bar:
            over.append((n, e, bar))
    errs = sorted(r[1] for r in ratios)
    print(f"{name} math={math}: gradient errors median {errs[len(errs) // 2]:.2e}, p90 {errs[int(0.9 * len(errs))]:.2e}, max {errs[-1]:.2e}; "
          f"{tight}/{len(errs)} parameters within max(2.5e-3, 4 x floor)")
    for r in sorted(ratios, key=lambda t: -t[1])[:3]:
        print(f"   err {r[1]:.2e} floor {r[2]:.2e} ({r[0]:.1f} x) {r[3]}")
    assert tight >= 0.9 * len(errs)
    assert len(over) <= 1, over
    if over:
        print(f"   over its bar (allowed: one, below twice the bar): {over[0][0]} err {over[0][1]:.2e} bar {over[0][2]:.2e}")
    print(f"{name} math={math}: log-probs err {lp_err:.2e}, losses {got}, worst gradient {worst[0]} at {worst[1]:.2f} of its bar")
    sd = m.state_dict()
    for k in [f for f in z.files if f.startswith("bn.")]:
        # running statistics after the step: batch means / variances of (deep) conv outputs, i.e. the forward's own tolerance
        assert torch.allclose(sd[k[3:]].cpu(), torch.from_numpy(z[k]), rtol=20 * tol, atol=20 * tol), k
    assert int(sd["backbone.bn1.num_batches_tracked"]) == 1


def test_frozen_backbone_recipe_and_optimizer(dev):
    """README's frozen recipe (``--freeze_layers 4`` -> dfs_freeze, slot_model.py:79-93): only the head gets gradients (no
    backbone backward is run), BatchNorm still updates its running statistics; then one step of ``torch.optim.AdamW`` (train.py:146)
    and of the fused ``adamw_step`` from the same state must agree, and the eval-mode forward must see the new weights."""
    from scouter_b200.train import adamw_step
    z, meta = load_golden("train_cfg2_resnest26d_224")
    m = build(meta, dev, L.MATH_TC)
    m.dfs_freeze(m.backbone, 4)
    assert not any(p.requires_grad for p in m.backbone.parameters())
    x = synth_images(meta["batch"], meta["cin"], meta["size"], meta["size"]).to(dev)
    tgt = torch.from_numpy(z["target"]).to(dev)
    rm0 = m.backbone.bn1.running_mean.clone()
    params = [p for p in m.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-3)
    out, (loss, nll, attn_loss) = m(x, tgt)
    loss.backward()
    names = meta["param_names"]
    pd = dict(m.named_parameters())
    scale = float(np.nanmax(z["grad_max"]))
    for i, n in enumerate(names):
        if n.startswith("backbone."):
            assert pd[n].grad is None, n
        elif not np.isnan(z["grad_max"][i]):
            den = max(float(z["grad_max"][i]), 1e-4 * scale)
            ref = torch.from_numpy(z["grad." + n])
            g = pd[n].grad.cpu()
            gg = g[:, ::8] if n == "conv1x1.weight" else g
            assert float((gg - ref).abs().max()) / den < max(1e-3, 8 * float(z["grad_floor"][i])), n
    assert not torch.equal(rm0, m.backbone.bn1.running_mean)
    # fused AdamW over flat buffers == torch.optim.AdamW (trainable, gradient-carrying parameters only)
    with_grad = [p for p in params if p.grad is not None]
    flat_p = torch.cat([p.detach().reshape(-1) for p in with_grad]).contiguous()
    flat_g = torch.cat([p.grad.reshape(-1) for p in with_grad]).contiguous()
    mm, vv = torch.zeros_like(flat_p), torch.zeros_like(flat_p)
    adamw_step(flat_p, flat_g, mm, vv, lr=1e-3, step=1)
    opt.step()
    torch.cuda.synchronize()
    ref_p = torch.cat([p.detach().reshape(-1) for p in with_grad])
    assert float((flat_p - ref_p).abs().max()) < 1e-6
    # the eval-mode forward now runs on the updated head and the updated running statistics
    m.eval()
    with torch.no_grad():
        e1 = m(x)
    m.train()
    out2, _ = m(x, tgt)
    assert torch.isfinite(e1).all() and not torch.equal(out2, out)


def test_train_mode_without_targets_and_no_grad(dev):
    z, meta = load_golden("train_cfg1_mnist_resnet18_260")
    m = build(meta, dev, L.MATH_TC)
    x = synth_images(meta["batch"], meta["cin"], meta["size"], meta["size"]).to(dev)
    out = m(x)                                               # train-mode forward, no labels: log-probs only (batch statistics)
    assert out.shape == (meta["batch"], 10) and torch.isfinite(out).all()
    with torch.no_grad():
        o2, losses = m(x, torch.from_numpy(z["target"]).to(dev))
    assert not losses[0].requires_grad
    with pytest.raises(IndexError):
        m(x, torch.full((meta["batch"],), 10, device=dev))
