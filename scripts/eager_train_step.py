"""Context for row f1: ONE training step (forward in .train(), loss.backward(), AdamW.step()) of (a) the UNMODIFIED reference
module under stock eager PyTorch and (b) this repo's SlotModel, on the same B200, same weights, same batch -- CUDA events.
Prints one JSON line.   python scripts/eager_train_step.py [--batch 32] [--steps 5]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scouter_b200 as sb  # noqa: E402
from baseline.refload import load_reference_model  # noqa: E402
from scouter_b200.synth import fill_state_dict, make_args  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=224)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
over = dict(model="resnest26d", num_classes=10, slots_per_class=1, power=2, to_k_layer=3, loss_status=-1, channel=2048)
ours = sb.SlotModel(make_args(**over))
sd = fill_state_dict(ours.state_dict(), seed=0)
ours.load_state_dict(sd)
ours = ours.to(dev).train()
x = torch.randn(a.batch, 3, a.size, a.size, device=dev)
y = torch.randint(0, 10, (a.batch,), device=dev)


def timed(model, steps, warmup=2):
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(warmup + steps):
        if i == warmup:
            torch.cuda.synchronize()
            ev[0].record()
        out, losses = model(x, y)
        opt.zero_grad(set_to_none=True)
        losses[0].backward()
        opt.step()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / steps, float(losses[0])


res = {"what": "one training step (train-mode forward + loss.backward() + AdamW.step()), resnest26d + negative xSlot, 1 GPU",
       "batch": a.batch, "size": a.size}
ms, loss = timed(ours, a.steps)
res["scouter_b200"] = {"ms_per_step": ms, "images_per_s": a.batch / ms * 1e3, "loss_after": loss}
ref = load_reference_model(state_dict=sd, feature_size=a.size // 32, **over)
if ref is None:
    res["reference_eager"] = {"unavailable": "no reference tree in baseline/_ref"}
else:
    ref = ref.to(dev).train()
    for name, tf32 in (("reference_eager_tf32", True), ("reference_eager_fp32", False)):
        saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        ref.load_state_dict(sd)
        ms, loss = timed(ref, a.steps)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
        res[name] = {"ms_per_step": ms, "images_per_s": a.batch / ms * 1e3, "loss_after": loss, "cudnn_allow_tf32": tf32}
print(json.dumps(res))
