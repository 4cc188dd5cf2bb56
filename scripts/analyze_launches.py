"""Join an ncu launch list (gpu__time_duration.sum per launch, one forward) with the lowered op program:
per-op shapes, algorithmic bytes/flops, achieved GB/s and TFLOP/s.  CPU-only (shape inference mirrors api.cu)."""
import argparse
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scouter_b200 as sb
from scouter_b200.synth import make_args
from scouter_b200 import _lib as L
from scouter_b200.plan import lower_backbone

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--size", type=int, default=224)
a = ap.parse_args()

rows = [r for r in csv.reader(open(a.csv)) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    name = r[ki].split("(")[0].split("::")[-1]
    launches.append((name, v))

m = sb.SlotModel(make_args(model="resnest26d", num_classes=10, slots_per_class=1, power=2, to_k_layer=3, loss_status=-1, channel=2048))
prog, feat = lower_backbone(m.backbone, L.MATH_TC)
shape = {0: (a.batch, a.size, a.size, 3)}
li = 0
tot = 0.0
print(f"{'op':12s} {'in':>18s} {'out':>18s} k s g {'us':>8s} {'GB/s':>7s} {'TF/s':>7s}  kernels")
for op in prog.ops:
    B, H, W, C = shape[op.src]
    k, s, p = op.kh, op.stride, op.pad
    kind = op.kind
    nk = 1
    if kind in (L.OP_STEM_CONV, L.OP_CONV):
        Ho, Wo, Co = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1, op.cout
        flops = 2.0 * B * Ho * Wo * Co * (C // op.groups) * k * k
        byts = 4.0 * (B * H * W * C + B * Ho * Wo * Co * (2 if op.src2 >= 0 else 1))
    elif kind == L.OP_MAXPOOL:
        Ho, Wo, Co = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1, C
        flops, byts = 0, 4.0 * B * (H * W + Ho * Wo) * C
    elif kind == L.OP_AVGPOOL:
        Ho, Wo, Co = -(-H // s), -(-W // s), C
        flops, byts = 0, 4.0 * B * (H * W + Ho * Wo) * C
    elif kind == L.OP_SPLAT_GAP:
        Ho, Wo, Co = 1, 1, op.cout
        flops, byts = 0, 4.0 * B * H * W * C
        # two kernels (partial + finish), or the finish kernel alone when the producing conv's epilogue wrote the partial sums
        nk = 2 if (li < len(launches) and "partial" in launches[li][0]) else 1
    elif kind == L.OP_SPLAT_APPLY:
        avd = bool(op.flags & L.F_AVD_POOL)
        Ho, Wo, Co = ((H + 1) // 2, (W + 1) // 2, op.cout) if avd else (H, W, op.cout)
        flops, byts = 0, 4.0 * B * (H * W * C + Ho * Wo * Co)
    else:
        Ho, Wo, Co = H, W, C
        flops = byts = 0
    shape[op.dst] = (B, Ho, Wo, Co)
    names, us = [], 0.0
    for _ in range(nk):
        if li < len(launches):
            names.append(launches[li][0][:26])
            us += launches[li][1]
            li += 1
    tot += us
    kn = {1: "stem", 2: "conv", 3: "maxpool", 4: "avgpool", 5: "splat_gap", 7: "splat_apply", 8: "gap", 9: "to_nchw"}.get(kind, str(kind))
    us_ = max(us, 1e-9)
    print(f"{kn:12s} {str((H, W, C)):>18s} {str((Ho, Wo, Co)):>18s} {k} {s} {op.groups} {us:8.1f} {byts / us_ / 1e3:7.0f} {flops / us_ / 1e6:7.1f}  {','.join(names)}")
print("backbone total us", round(tot), "| remaining launches:", [(n[:24], round(v, 1)) for n, v in launches[li:]])
