"""Golden vectors for row f1 (training step), dumped from the UNMODIFIED reference in the build container.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_train [--check]     # writes tests/golden/train_*.npz

Per case: the reference ``SlotModel`` (``oracle/refshim.py``) with the deterministic synthetic weights in ``.train()``
mode runs ``engine.py:28-33`` once -- ``outputs, (loss, nll, attn_loss) = model(x, target)``; ``loss.backward()`` --
and the file keeps: the log-probs, the three losses, max|grad| and the L2 norm of every parameter's gradient, its fp32-vs-fp64 noise floor (in
``named_parameters`` order; NaN for parameters the graph does not reach: ``slot.to_q.*``), the full gradients of the
head parameters and of a few backbone tensors, and the BatchNorm running statistics after the step for the first and
last BatchNorm.  TEST INFRASTRUCTURE ONLY: nothing in scouter_b200 implements the backward yet; these files and
``oracle/train.py`` are the bar it will be held to.
"""
from __future__ import annotations

import argparse
import json
import os

import numpy as np
import torch

from scouter_b200.synth import fill_state_dict, synth_images, synth_labels

from . import refshim
from .train import train_step

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    "train_cfg2_resnest26d_224": dict(args=dict(model="resnest26d", dataset="ImageNet", channel=2048, num_classes=10,
                                                slots_per_class=1, power=2, to_k_layer=3, loss_status=1, lambda_value=1.0),
                                      batch=4, cin=3, size=224),
    "train_cfg3_resnest26d_neg_224": dict(args=dict(model="resnest26d", dataset="ImageNet", channel=2048, num_classes=10,
                                                    slots_per_class=1, power=2, to_k_layer=3, loss_status=-1, lambda_value=1.0),
                                          batch=3, cin=3, size=224),
    "train_cfg1_mnist_resnet18_260": dict(args=dict(model="resnet18", dataset="MNIST", channel=512, num_classes=10,
                                                    slots_per_class=1, power=1, to_k_layer=1, loss_status=1, lambda_value=1.0),
                                          batch=4, cin=1, size=260),
}
# gradients kept in full (besides every non-backbone parameter)
FULL = ("backbone.conv1.0.weight", "backbone.conv1.weight", "backbone.bn1.weight", "backbone.bn1.bias",
        "backbone.layer1.0.conv2.fc2.bias", "backbone.layer4.1.bn3.weight", "backbone.layer4.1.bn2.weight")


def run_case(name, case, check):
    a = case["args"]
    torch.manual_seed(0)
    ref = refshim.reference_slot_model(**a)
    sd = fill_state_dict(ref.state_dict(), seed=0)
    ref.load_state_dict(sd, strict=True)
    ref.feature_size = {260: 9, 224: 7}[case["size"]]
    ref.train()
    x = synth_images(case["batch"], case["cin"], case["size"], case["size"])
    tgt = synth_labels(case["batch"], a["num_classes"])
    out, (loss, nll, attn_loss) = ref(x, tgt)
    loss.backward()
    names = [n for n, _ in ref.named_parameters()]
    gmax = np.array([float("nan") if p.grad is None else float(p.grad.abs().max()) for _, p in ref.named_parameters()])
    gl2 = np.array([float("nan") if p.grad is None else float(p.grad.double().norm()) for _, p in ref.named_parameters()])
    z = {"log_probs": out.detach().numpy(), "losses": np.array([float(loss), float(nll), float(attn_loss)]),
         "grad_max": gmax, "grad_l2": gl2, "target": tgt.numpy()}
    for n, p in ref.named_parameters():
        if p.grad is not None and (not n.startswith("backbone.") or n in FULL):
            # conv1x1.weight (64 x ch x 1 x 1) is kept for every 8th input channel only (file size)
            z["grad." + n] = p.grad.numpy()[:, ::8] if n == "conv1x1.weight" else p.grad.numpy()
    # the reference's own fp32 noise on these gradients: same step in fp64 (the sum-normalisation of
    # slot_attention.py:56 amplifies rounding noise, so the bar for a parameter is max(1e-3, 4 x this floor))
    ref64 = refshim.reference_slot_model(**a).double()
    ref64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()})
    ref64.feature_size = ref.feature_size
    ref64.train()
    loss64 = ref64(x.double(), tgt)[1][0]
    loss64.backward()
    scale = np.nanmax(gmax)
    floor = np.array([float("nan") if p.grad is None else
                      float((p.grad.double() - q.grad).abs().max()) / max(float(q.grad.abs().max()), 1e-4 * scale)
                      for (_, p), (_, q) in zip(ref.named_parameters(), ref64.named_parameters())])
    z["grad_floor"] = floor
    z["loss64"] = np.array(float(loss64))
    after = ref.state_dict()
    bn_keys = [k for k in after if k.endswith(("running_mean", "running_var"))]
    for k in bn_keys[:2] + bn_keys[-2:]:
        z["bn." + k] = after[k].numpy()
    z["meta"] = np.array(json.dumps({"args": a, "batch": case["batch"], "cin": case["cin"], "size": case["size"],
                                     "param_names": names}))
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **z)

    o = train_step(a["model"], sd, x, tgt, num_classes=a["num_classes"], slots_per_class=a["slots_per_class"],
                   loss_status=a["loss_status"], power=a["power"], lambda_value=a["lambda_value"])
    worst = 0.0
    for (n, p), fl in zip(ref.named_parameters(), floor):
        g = o["grads"][n]
        if p.grad is None:
            assert g is None, n
            continue
        e = float((g - p.grad).abs().max()) / max(float(p.grad.abs().max()), 1e-4 * scale)
        worst = max(worst, e / max(1e-3, 4 * fl))
    bn = max(float((v - after[k]).abs().max()) for k, v in o["bn_updates"].items())
    print(f"[{name}] loss {float(loss):.6f} max|grad| {scale:.3e}; oracle-vs-reference: loss {abs(float(o['loss'] - loss)):.1e} "
          f"grads {worst:.2f} of the bar max(1e-3, 4 x fp64 floor; worst floor {np.nanmax(floor):.1e}) bn {bn:.1e}")
    if check:
        assert abs(float(o["loss"] - loss)) < 1e-5 and worst < 1.0 and bn < 1e-6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    for name, case in CASES.items():
        run_case(name, case, a.check)


if __name__ == "__main__":
    main()
