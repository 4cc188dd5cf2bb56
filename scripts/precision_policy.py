"""Precision-policy experiment: which backbone stages can run as ONE tf32 pass on rounded operands
(SCOUTER_F_TF32_1PASS) while the rest keeps the error-compensated product?  For every policy: the parity metrics of
tests/test_gpu_parity.py::test_slot_model_vs_reference_golden on every model golden (log-prob error against north_star's
1e-3, attention-map error against max(1e-3, 8 x reference floor)), and the forward throughput of cfg 3 at B=256.
Writes a table to stdout (kept as profiles/r02_precision_policy.txt).   python scripts/precision_policy.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_names, load_golden  # noqa: E402
import scouter_b200 as sb  # noqa: E402
from scouter_b200 import _lib as L  # noqa: E402
from scouter_b200.synth import fill_state_dict, make_args, synth_images  # noqa: E402

ATTN_TOL_TC = {"f4_resnest50d_224": 5e-3}    # the recorded exception of tests/test_gpu_parity.py (resnest50d attention maps)
POLICIES = [(), ("stem",), ("stem", "layer1"), ("stem", "layer1", "layer2"), ("stem", "layer1", "layer2", "layer3"),
            ("stem", "layer1", "layer2", "layer3", "layer4"), ("layer1",), ("layer2",), ("layer3",), ("layer4",)]


def scaled_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / max(1.0, float(b.abs().max())))


def parity(name, policy, dev):
    z, meta = load_golden(name)
    m = sb.SlotModel(make_args(**meta["args"]))
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=0))
    m = m.to(dev).eval()
    m.math, m.fast_stages, m.keep_attn = L.MATH_TC, policy, True
    x = synth_images(meta["batch"], meta["cin"], meta["size"], meta["size"]).to(dev)
    with torch.no_grad():
        out = m(x)
    floor = scaled_err(z["log_probs"], z["log_probs64"])
    afloor = float(np.abs(z["attn"].astype(np.float64) - z["attn64"]).max())
    e_lp = scaled_err(out, z["log_probs"])
    e_at = float((m.last_attn.cpu() - torch.from_numpy(z["attn"])).abs().max())
    return e_lp, max(1e-3, 4 * floor), e_at, max(1e-3, 8 * afloor, ATTN_TOL_TC.get(name, 0.0))


def throughput(policy, dev, batch=256, iters=10):
    m = sb.SlotModel(make_args(model="resnest26d", num_classes=10, slots_per_class=1, power=2, to_k_layer=3, loss_status=1, channel=2048))
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=0))
    m = m.to(dev).eval()
    m.math, m.fast_stages = L.MATH_TC, policy
    x = torch.randn(batch, 3, 224, 224, device=dev)
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            m(x)
        e1.record()
        torch.cuda.synchronize()
    return batch * iters / (e0.elapsed_time(e1) * 1e-3)


def main():
    dev = torch.device("cuda", 0)
    names = golden_names("model")
    print("# error-compensated tensor-core mode (SCOUTER_MATH=tc) with the listed stages run as single-pass tf32 (SCOUTER_TC_FAST_STAGES)")
    print("# per golden: log-prob err / its bar max(1e-3, 4 x floor), attention err / its bar max(1e-3, 8 x floor); '!' = over the bar,")
    print("# '~' = inside the bar with less than 2x margin.  img/s: cfg 3 forward, B=256, 224^2, eager calls, CUDA events.")
    for pol in POLICIES:
        ips = throughput(pol, dev)
        cells, worst = [], 0.0
        for n in names:
            e_lp, t_lp, e_at, t_at = parity(n, pol, dev)
            worst = max(worst, e_lp / t_lp, e_at / t_at)
            mark = lambda e, t: "!" if e >= t else ("~" if 2 * e >= t else " ")
            cells.append(f"{n.split('_')[0]}: lp {e_lp:.1e}/{t_lp:.0e}{mark(e_lp, t_lp)} at {e_at:.1e}/{t_at:.0e}{mark(e_at, t_at)}")
        verdict = "FAIL" if worst >= 1 else ("marginal" if worst >= 0.5 else "ok (>=2x margin)")
        print(f"{'+'.join(pol) or 'none (shipped default)':42s} {ips:8.0f} img/s  worst err/bar {worst:5.2f}  {verdict}")
        for c in cells:
            print("      " + c)
        sys.stdout.flush()


if __name__ == "__main__":
    main()
