"""CPU: the training-step op program of ``scouter_b200.plan`` (``lower_backbone_train`` + ``backward_schedule``; row f1
groundwork, host side only) interpreted op by op -- forward in torch, every backward op through the host emulation of
the draft CUDA kernel bodies (tests/draft_emu.py) -- and compared with the train-mode oracle (the reference's
``loss.backward()``).  This is the executor loop the C++ side will run, including the gradient-accumulation flags."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import draft_emu as E
import scouter_b200 as sb
from oracle import head as oh
from oracle.train import train_step
from scouter_b200.plan import backward_schedule, lower_backbone_train
from scouter_b200.synth import fill_state_dict, make_args, synth_images, synth_labels
from test_train_step_draft_composition import to_nchw, to_nhwc


def run_forward(ops, sd, x_nhwc):
    """-> (buffers, per-op saved state).  Values come from the torch ops the oracle runs (see the composition test for why);
    the emulated BatchNorm forward supplies the saved statistics and the running-statistic updates."""
    buf, saved, bn_after = {0: x_nhwc}, {}, {}
    for i, op in enumerate(ops):
        k = op["kind"]
        if k == "conv":
            w, b = sd[op["key"] + ".weight"], sd.get(op["key"] + ".bias")
            buf[op["dst"]] = to_nhwc(F.conv2d(to_nchw(buf[op["src"]]), w, b, op["stride"], op["pad"], 1, op["groups"]))
        elif k == "bn":
            key, x = op["key"], buf[op["src"]]
            res = buf[op["res"]] if op["res"] >= 0 else None
            rm, rv = sd[key + ".running_mean"].numpy().copy(), sd[key + ".running_var"].numpy().copy()
            _, mean, rstd = E.bn_train_forward(x, sd[key + ".weight"].numpy(), sd[key + ".bias"].numpy(), rm, rv, residual=res, relu=op["relu"])
            bn_after[key + ".running_mean"], bn_after[key + ".running_var"] = rm, rv
            yt = F.batch_norm(to_nchw(x), None, None, sd[key + ".weight"], sd[key + ".bias"], True, 0.1, 1e-5)
            if res is not None:
                yt = yt + to_nchw(res)
            buf[op["dst"]] = to_nhwc(torch.relu(yt) if op["relu"] else yt)
            saved[i] = (mean, rstd)
        elif k == "maxpool":
            buf[op["dst"]] = to_nhwc(F.max_pool2d(to_nchw(buf[op["src"]]), 3, 2, 1))
        elif k == "avgpool2":
            buf[op["dst"]] = to_nhwc(F.avg_pool2d(to_nchw(buf[op["src"]]), 2, 2, ceil_mode=True, count_include_pad=False))
        elif k == "avgpool3":
            buf[op["dst"]] = to_nhwc(F.avg_pool2d(to_nchw(buf[op["src"]]), 3, 2, 1))
        elif k == "splat_gap":
            x2 = buf[op["src"]]
            c = x2.shape[-1] // 2
            buf[op["dst"]] = (x2[..., :c] + x2[..., c:]).mean((1, 2), keepdims=True).astype(np.float32)
        elif k == "splat_mix":
            x2, lg = buf[op["src"]], buf[op["logits"]]
            b, c = x2.shape[0], x2.shape[-1] // 2
            att = torch.softmax(torch.from_numpy(lg).reshape(b, 2, c), dim=1).numpy()
            buf[op["dst"]] = x2[..., :c] * att[:, None, None, 0] + x2[..., c:] * att[:, None, None, 1]
            saved[op["dst"]] = att
        else:
            raise AssertionError(k)
    return buf, saved, bn_after


def run_backward(ops, sched, sd, buf, saved, d_feat, feat):
    d, grads = {feat: d_feat}, {}

    def put(e, b, g):
        assert e["acc"][b] == (b in d), (e["kind"], b)       # the schedule's accumulate flag == "someone wrote this gradient already"
        d[b] = d[b] + g if e["acc"][b] else g

    for e in sched:
        k = e["kind"]
        if k == "conv":
            w = sd[e["key"] + ".weight"]
            has_b = (e["key"] + ".bias") in sd
            w_ohwi = np.ascontiguousarray(w.permute(0, 2, 3, 1).numpy())
            dx, dw, db = E.conv_backward(buf[e["src"]], d[e["dst"]], w_ohwi, e["stride"], e["pad"], e["groups"], bias=has_b, need_dx=e["need_dx"])
            grads[e["key"] + ".weight"] = torch.from_numpy(np.ascontiguousarray(dw.transpose(0, 3, 1, 2)))
            if has_b:
                grads[e["key"] + ".bias"] = torch.from_numpy(db)
            if e["need_dx"]:
                put(e, e["src"], dx)
        elif k == "bn":
            i = next(j for j, op in enumerate(ops) if op["kind"] == "bn" and op["dst"] == e["dst"])
            mean, rstd = saved[i]
            dx, dg, db, dres = E.bn_train_backward(buf[e["src"]], buf[e["dst"]], d[e["dst"]], sd[e["key"] + ".weight"].numpy(), mean, rstd,
                                                   relu=e["relu"], want_residual=e["res"] >= 0)
            grads[e["key"] + ".weight"], grads[e["key"] + ".bias"] = torch.from_numpy(dg), torch.from_numpy(db)
            put(e, e["src"], dx)
            if e["res"] >= 0:
                put(e, e["res"], dres)
        elif k in ("maxpool", "avgpool2", "avgpool3"):
            put(e, e["src"], E.pool_backward({"maxpool": 0, "avgpool2": 1, "avgpool3": 2}[k], buf[e["src"]], d[e["dst"]]))
        elif k == "splat_mix":
            att = saved[e["dst"]]
            b, c = att.shape[0], att.shape[2]
            put(e, e["logits"], E.splat_backward_logits(buf[e["src"]], d[e["dst"]], att).reshape(b, 1, 1, 2 * c))
        elif k == "splat_gap":
            att = saved[e["d_mix"]]
            b, c = att.shape[0], att.shape[2]
            put(e, e["src"], E.splat_backward_apply(buf[e["src"]], d[e["d_mix"]], att, d[e["dst"]].reshape(b, c)))
    return grads


def program_gradients(m, a, sd, x, tgt):
    """One training-step forward + backward of SlotModel ``m`` through the op program and the emulated draft kernels.
    -> ({state_dict key: gradient tensor}, {running-statistic key: value after the step})"""
    ops, nbuf, feat = lower_backbone_train(m.backbone)
    sched = backward_schedule(ops, feat)
    assert len(sched) == len(ops) and sorted(o["dst"] for o in ops) == list(range(1, nbuf))
    buf, saved, bn_after = run_forward(ops, sd, to_nhwc(x))
    h = buf[feat]
    bsz, fh, fw, ch = h.shape
    n, S, L = fh * fw, 10, a["to_k_layer"]
    with torch.no_grad():
        ho = oh.head_forward(sd, to_nchw(h), num_classes=10, slots_per_class=1, loss_status=1, power=a["power"], return_attn=True)
    g_logits = (ho["log_probs"].exp() - F.one_hot(tgt, 10)) / bsz
    mean_attn = float(ho["attn"].sum()) / (bsz * S * n)
    coef = 1.0 * a["power"] * mean_attn ** (a["power"] - 1) / (bsz * S * n)
    pe = oh.sine_pe(64, fh, fw).reshape(64, n).t().numpy()
    d_feat, grads = E.head_backward(h.reshape(bsz, n, ch), {k: v.numpy() for k, v in sd.items() if not k.startswith("backbone.")}, pe,
                                    g_logits.numpy(), coef, 10, 1, 1, L)
    grads = {k: torch.from_numpy(np.ascontiguousarray(v)).reshape(sd[k].shape) for k, v in grads.items()}
    grads.update(run_backward(ops, sched, sd, buf, saved, d_feat.reshape(bsz, fh, fw, ch), feat))
    return grads, bn_after


@pytest.mark.timeout(900)
@pytest.mark.parametrize("model,extra", [("resnest26d", dict()), ("resnet18", dict(dataset="MNIST", channel=512, to_k_layer=1, power=1))])
def test_train_program_interpreted_matches_train_oracle(model, extra, min_bar=5e-4):
    if E.lib() is None:
        pytest.skip("g++ not available")
    a = dict(model=model, num_classes=10, slots_per_class=1, power=2, to_k_layer=3, loss_status=1, lambda_value=1.0)
    a.update(extra)
    m = sb.SlotModel(make_args(**a))
    sd = fill_state_dict(m.state_dict(), seed=0)
    B, size, cin = 3, 64, (1 if model == "resnet18" else 3)
    x, tgt = synth_images(B, cin, size, size), synth_labels(B, 10)
    kw = dict(num_classes=10, slots_per_class=1, loss_status=1, power=a["power"], lambda_value=1.0)
    ref, ref64 = train_step(model, sd, x, tgt, **kw), train_step(model, sd, x, tgt, dtype=torch.float64, **kw)
    grads, bn_after = program_gradients(m, a, sd, x, tgt)

    scale = max(float(g.abs().max()) for g in ref["grads"].values() if g is not None)
    checked, bad = 0, []
    for k, g_ref in ref["grads"].items():
        if g_ref is None:
            continue
        got = grads[k].reshape(g_ref.shape)
        if k.endswith(".conv2.fc1.bias"):                     # exactly zero in exact arithmetic: rounding noise on both sides
            assert float(got.abs().max()) < 1e-3 * scale
        else:
            den = max(float(g_ref.abs().max()), 1e-4 * scale)
            floor = float((g_ref.double() - ref64["grads"][k]).abs().max()) / den
            err = float((got - g_ref).abs().max()) / den
            if not err < max(min_bar, 8 * floor):
                bad.append((k, err, floor))
        checked += 1
    assert not bad, sorted(bad, key=lambda t: -t[1])[:12]
    assert checked == len(grads)
    for k, v in ref["bn_updates"].items():
        assert np.allclose(bn_after[k], v.numpy(), rtol=1e-4, atol=1e-5), k
