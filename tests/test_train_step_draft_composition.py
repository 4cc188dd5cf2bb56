"""CPU: the row-f1 DRAFT kernels composed into one full training-step backward (resnest14d / resnest26d + xSlot, tiny images) and
compared with the train-mode oracle (oracle/train.py = the reference's ``loss.backward()``).

The forward runs in torch (it is not under test); every backward op -- BatchNorm (train statistics), conv data / weight
gradients, max / average pools, split attention, the xSlot head -- goes through the host emulation of the draft CUDA
kernel bodies (tests/draft_emu.py).  This checks the *composition*: which gradient flows where, in the order the backward
op program of the library will run it.  Nothing here has run on a GPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import draft_emu as E
import scouter_b200 as sb
from oracle import head as oh
from oracle.train import train_step
from scouter_b200.plan import dgrad_weights
from scouter_b200.synth import fill_state_dict, make_args, synth_images, synth_labels


def to_nhwc(t):
    return np.ascontiguousarray(t.permute(0, 2, 3, 1).numpy())


def to_nchw(a):
    return torch.from_numpy(np.ascontiguousarray(a)).permute(0, 3, 1, 2).contiguous()     # same memory format as the oracle's tensors


class Tape:
    """Forward in torch on NHWC numpy arrays; every op pushes a closure that maps d_out -> d_in and records parameter
    gradients under the reference's state_dict keys."""

    def __init__(self, sd):
        self.sd = sd
        self.grads = {}
        self.bn_after = {}

    def _acc(self, key, g):
        g = torch.from_numpy(np.ascontiguousarray(g)).reshape(self.sd[key].shape)
        self.grads[key] = self.grads.get(key, 0) + g

    def conv(self, x, key, stride=1, pad=0, groups=1, need_dx=True):
        w = self.sd[key + ".weight"]
        bias = self.sd.get(key + ".bias")
        y = to_nhwc(F.conv2d(to_nchw(x), w, bias, stride, pad, 1, groups))
        w_ohwi = np.ascontiguousarray(w.permute(0, 2, 3, 1).numpy())

        def back(dy):
            dx, dw, db = E.conv_backward(x, dy, w_ohwi, stride, pad, groups, bias=bias is not None, need_dx=need_dx)
            if need_dx and stride == 1:
                # the library's route for stride-1 convs: the data gradient as a FORWARD conv with plan.dgrad_weights
                k = w_ohwi.shape[1]
                wd = dgrad_weights(torch.from_numpy(w_ohwi), groups)                       # (Cin, k, k, Cout/g)
                via_fwd = to_nhwc(F.conv2d(to_nchw(dy), wd.permute(0, 3, 1, 2).contiguous(), None, 1, k - 1 - pad, 1, groups))
                assert np.abs(via_fwd - dx).max() <= 1e-5 * max(1.0, np.abs(dx).max()), key
            self._acc(key + ".weight", np.ascontiguousarray(dw.transpose(0, 3, 1, 2)))
            if bias is not None:
                self._acc(key + ".bias", db)
            return dx
        return y, back

    def bn(self, x, key, relu=True, residual=None):
        g, b = self.sd[key + ".weight"].numpy(), self.sd[key + ".bias"].numpy()
        rm, rv = self.sd[key + ".running_mean"].numpy().copy(), self.sd[key + ".running_var"].numpy().copy()
        y_emu, mean, rstd = E.bn_train_forward(x, g, b, rm, rv, residual=residual, relu=relu)
        self.bn_after[key + ".running_mean"], self.bn_after[key + ".running_var"] = rm, rv
        # Downstream values (and with them every ReLU mask) come from the same torch op the oracle runs, so that the
        # comparison below sees rounding only: with 3 images per batch a 1e-6 difference in one BatchNorm grows to 3e-5
        # four blocks later and flips a ReLU mask here and there, which moves whole gradient tensors by 1e-2.  The
        # emulated forward kernel is held to the torch result right here instead.
        yt = F.batch_norm(to_nchw(x), None, None, self.sd[key + ".weight"], self.sd[key + ".bias"], True, 0.1, 1e-5)
        if residual is not None:
            yt = yt + to_nchw(residual)
        y = to_nhwc(torch.relu(yt) if relu else yt)
        assert np.abs(y_emu - y).max() <= 1e-4 * max(1.0, np.abs(y).max()), key

        def back(dy):
            dx, dg, db, dres = E.bn_train_backward(x, y, dy, g, mean, rstd, relu=relu, want_residual=residual is not None)
            self._acc(key + ".weight", dg)
            self._acc(key + ".bias", db)
            return dx, dres
        return y, back


def conv_bn(t, x, ckey, bkey, stride=1, pad=0, groups=1, relu=True, residual=None, need_dx=True):
    y0, cb = t.conv(x, ckey, stride, pad, groups, need_dx)
    y, bb = t.bn(y0, bkey, relu, residual)

    def back(dy):
        d0, dres = bb(dy)
        return cb(d0), dres
    return y, back


def resnest_block(t, x, p, avd, down_pool, has_down=True):
    o1, b1 = conv_bn(t, x, p + ".conv1", p + ".bn1")
    x2, b2 = conv_bn(t, o1, p + ".conv2.conv", p + ".conv2.bn0", pad=1, groups=2)
    bsz, h, w, c2 = x2.shape
    c = c2 // 2
    gap = (x2[..., :c] + x2[..., c:]).mean((1, 2), keepdims=True).astype(np.float32)                 # (B,1,1,C)
    a1, bf1 = conv_bn(t, gap, p + ".conv2.fc1", p + ".conv2.bn1")
    logits, bf2 = t.conv(a1, p + ".conv2.fc2")
    att = torch.softmax(torch.from_numpy(logits).reshape(bsz, 2, c), dim=1).numpy()
    o2 = x2[..., :c] * att[:, None, None, 0] + x2[..., c:] * att[:, None, None, 1]
    pooled = to_nhwc(F.avg_pool2d(to_nchw(o2), 3, 2, 1)) if avd else o2
    res_in = to_nhwc(F.avg_pool2d(to_nchw(x), 2, 2, ceil_mode=True, count_include_pad=False)) if down_pool else x
    if has_down:
        res, bd = conv_bn(t, res_in, p + ".downsample.1", p + ".downsample.2", relu=False)
    else:                                                                                          # identity shortcut
        res, bd = x, (lambda d_res: (d_res, None))
    out, b3 = conv_bn(t, pooled, p + ".conv3", p + ".bn3", relu=True, residual=res)

    def back(d_out):
        d_pooled, d_res = b3(d_out)
        d_o2 = E.pool_backward(2, o2, d_pooled) if avd else d_pooled
        d_logit = E.splat_backward_logits(x2, d_o2, att)                                               # (B,2,C) radix-major
        d_a1 = bf2(d_logit.reshape(bsz, 1, 1, 2 * c))
        d_gap, _ = bf1(d_a1)
        d_x2 = E.splat_backward_apply(x2, d_o2, att, d_gap.reshape(bsz, c))
        d_o1, _ = b2(d_x2)
        d_x, _ = b1(d_o1)
        d_res_in, _ = bd(d_res)
        return d_x + (E.pool_backward(1, x, d_res_in) if down_pool else d_res_in)
    return out, back


@pytest.mark.timeout(900)
@pytest.mark.parametrize("model,per_layer", [("resnest14d", 1), ("resnest26d", 2)])
def test_full_backward_composed_from_draft_kernels_matches_train_oracle(model, per_layer):
    if E.lib() is None:
        pytest.skip("g++ not available")
    args = dict(model=model, num_classes=10, slots_per_class=1, power=2, to_k_layer=3, loss_status=1, lambda_value=1.0)
    m = sb.SlotModel(make_args(**args))
    sd = fill_state_dict(m.state_dict(), seed=0)
    B, size = 3, 64
    x = synth_images(B, 3, size, size)
    tgt = synth_labels(B, 10)
    ref = train_step(model, sd, x, tgt, num_classes=10, slots_per_class=1, loss_status=1, power=2, lambda_value=1.0)
    ref64 = train_step(model, sd, x, tgt, num_classes=10, slots_per_class=1, loss_status=1, power=2, lambda_value=1.0,
                       dtype=torch.float64)

    # ---- forward with a tape -------------------------------------------------------------------------------------------
    t = Tape(sd)
    backs = []
    h, bk = conv_bn(t, to_nhwc(x), "backbone.conv1.0", "backbone.conv1.1", stride=2, pad=1, need_dx=False); backs.append(bk)
    h, bk = conv_bn(t, h, "backbone.conv1.3", "backbone.conv1.4", pad=1); backs.append(bk)
    h, bk = conv_bn(t, h, "backbone.conv1.6", "backbone.bn1", pad=1); backs.append(bk)
    pool_in = h
    h = to_nhwc(F.max_pool2d(to_nchw(h), 3, 2, 1))
    blocks = []
    for li in range(1, 5):
        for bi in range(per_layer):
            first = bi == 0
            h, bk = resnest_block(t, h, f"backbone.layer{li}.{bi}", avd=first and li > 1, down_pool=first and li > 1, has_down=first)
            blocks.append(bk)
    bsz, fh, fw, ch = h.shape
    feat_tokens = h.reshape(bsz, fh * fw, ch)
    with torch.no_grad():
        ho = oh.head_forward(sd, to_nchw(h), num_classes=10, slots_per_class=1, loss_status=1, power=2, return_attn=True)
    assert float((ho["log_probs"] - ref["log_probs"]).abs().max()) < 1e-3          # the taped forward is the oracle's forward
    S, n = 10, fh * fw
    g_logits = (ho["log_probs"].exp() - F.one_hot(tgt, 10)) / bsz                   # d nll / d logits
    mean_attn = float(ho["attn"].sum()) / (bsz * S * n)
    coef = 1.0 * 2 * mean_attn ** (2 - 1) / (bsz * S * n)                           # lambda * power * m^(power-1) / (B S n)
    pe = oh.sine_pe(64, fh, fw).reshape(64, n).t().numpy()

    # ---- backward: head, blocks in reverse, max-pool, stem ---------------------------------------------------------------
    d_feat, head_grads = E.head_backward(feat_tokens, {k: v.numpy() if k != "slot.initial_slots" else v.numpy() for k, v in sd.items()
                                                       if not k.startswith("backbone.")}, pe, g_logits.numpy(), coef, 10, 1, 1, 3)
    for k, v in head_grads.items():
        t._acc(k, v)
    d = d_feat.reshape(bsz, fh, fw, ch)
    for bk in reversed(blocks):
        d = bk(d)
    d = E.pool_backward(0, pool_in, d)
    for bk in reversed(backs):
        d, _ = bk(d)

    # ---- compare every parameter gradient and the running statistics ---------------------------------------------------
    scale = max(float(g.abs().max()) for g in ref["grads"].values() if g is not None)
    checked = 0
    for k, g_ref in ref["grads"].items():
        if g_ref is None:
            assert k.startswith("slot.to_q.") and k not in t.grads
            continue
        if k.endswith(".conv2.fc1.bias"):
            # a bias in front of a train-mode BatchNorm has an exactly zero gradient; both sides hold rounding noise only
            assert float(t.grads[k].abs().max()) < 1e-3 * scale and float(g_ref.abs().max()) < 1e-3 * scale, k
            checked += 1
            continue
        den = max(float(g_ref.abs().max()), 1e-4 * scale)
        floor = float((g_ref.double() - ref64["grads"][k]).abs().max()) / den          # the oracle's own fp32 noise on this tensor
        err = float((t.grads[k] - g_ref).abs().max()) / den
        assert err < max(5e-4, 8 * floor), (k, err, floor)      # measured: <= 2e-4
        checked += 1
    assert checked == len(t.grads) == sum(g is not None for g in ref["grads"].values())
    for k, v in ref["bn_updates"].items():
        assert np.allclose(t.bn_after[k], v.numpy(), rtol=1e-4, atol=1e-5), k


def basic_block(t, x, p, stride, has_down):
    o1, b1 = conv_bn(t, x, p + ".conv1", p + ".bn1", stride=stride, pad=1)
    if has_down:                                                                                   # 1x1 stride-s conv + BN shortcut
        res, bd = conv_bn(t, x, p + ".downsample.0", p + ".downsample.1", stride=stride, relu=False)
    else:
        res, bd = x, (lambda d_res: (d_res, None))
    out, b2 = conv_bn(t, o1, p + ".conv2", p + ".bn2", pad=1, relu=True, residual=res)

    def back(d_out):
        d_o1, d_res = b2(d_out)
        d_x, _ = b1(d_o1)
        d_sc, _ = bd(d_res)
        return d_x + d_sc
    return out, back


@pytest.mark.timeout(900)
def test_full_backward_resnet18_mnist_from_draft_kernels_matches_train_oracle():
    """cfg 1 (MNIST resnet18 + xSlot, 1-channel 3x3 s2 stem, slot_model.py:23-24): strided 3x3 convs and conv shortcuts go
    through the general-stride data-gradient gather."""
    if E.lib() is None:
        pytest.skip("g++ not available")
    args = dict(model="resnet18", dataset="MNIST", channel=512, num_classes=10, slots_per_class=1, power=1, to_k_layer=1,
                loss_status=1, lambda_value=1.0)
    m = sb.SlotModel(make_args(**args))
    sd = fill_state_dict(m.state_dict(), seed=0)
    B, size = 3, 64
    x = synth_images(B, 1, size, size)
    tgt = synth_labels(B, 10)
    kw = dict(num_classes=10, slots_per_class=1, loss_status=1, power=1, lambda_value=1.0)
    ref = train_step("resnet18", sd, x, tgt, **kw)
    ref64 = train_step("resnet18", sd, x, tgt, dtype=torch.float64, **kw)
    t = Tape(sd)
    h, stem_back = conv_bn(t, to_nhwc(x), "backbone.conv1", "backbone.bn1", stride=2, pad=1, need_dx=False)
    pool_in = h
    h = to_nhwc(F.max_pool2d(to_nchw(h), 3, 2, 1))
    blocks = []
    for li in range(1, 5):
        for bi in range(2):
            first = bi == 0 and li > 1
            h, bk = basic_block(t, h, f"backbone.layer{li}.{bi}", 2 if first else 1, first)
            blocks.append(bk)
    bsz, fh, fw, ch = h.shape
    n, S = fh * fw, 10
    with torch.no_grad():
        ho = oh.head_forward(sd, to_nchw(h), num_classes=10, slots_per_class=1, loss_status=1, power=1, return_attn=True)
    g_logits = (ho["log_probs"].exp() - F.one_hot(tgt, 10)) / bsz
    coef = 1.0 * 1 / (bsz * S * n)                                                                 # power 1: lambda / (B S n)
    pe = oh.sine_pe(64, fh, fw).reshape(64, n).t().numpy()
    d_feat, head_grads = E.head_backward(h.reshape(bsz, n, ch), {k: v.numpy() for k, v in sd.items() if not k.startswith("backbone.")},
                                         pe, g_logits.numpy(), coef, 10, 1, 1, 1)
    for k, v in head_grads.items():
        t._acc(k, v)
    d = d_feat.reshape(bsz, fh, fw, ch)
    for bk in reversed(blocks):
        d = bk(d)
    d = E.pool_backward(0, pool_in, d)
    stem_back(d)
    scale = max(float(g.abs().max()) for g in ref["grads"].values() if g is not None)
    checked = 0
    for k, g_ref in ref["grads"].items():
        if g_ref is None:
            continue
        den = max(float(g_ref.abs().max()), 1e-4 * scale)
        floor = float((g_ref.double() - ref64["grads"][k]).abs().max()) / den
        err = float((t.grads[k] - g_ref).abs().max()) / den
        assert err < max(5e-4, 8 * floor), (k, err, floor)
        checked += 1
    assert checked == len(t.grads)
    for k, v in ref["bn_updates"].items():
        assert np.allclose(t.bn_after[k], v.numpy(), rtol=1e-4, atol=1e-5), k
