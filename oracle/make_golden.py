"""Oracle (test infrastructure): generate ``tests/golden/*.npz`` from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden            # writes tests/golden/
    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden --check    # also asserts oracle == reference

For every case it
  1. builds the reference module (``sloter.slot_model.SlotModel`` / ``sloter.utils.slot_attention.SlotAttention``)
     through ``oracle/refshim.py``,
  2. loads the deterministic synthetic ``state_dict`` of ``scouter_b200.synth`` (strict -- this also pins the
     key/shape contract of SURVEY.md App. D),
  3. runs the reference forward in fp32 (the golden) and in fp64 (its own noise floor),
  4. runs the oracle restatement (``oracle/head.py``, ``oracle/backbone.py``) on the same inputs and records
     the difference (asserted small with --check and again by tests/test_oracle_golden.py).

Weights/inputs are not stored: they are regenerated bit-identically from (name, shape, seed) on any box.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

from scouter_b200.synth import fill_state_dict, synth_images, synth_labels

from . import backbone as ob
from . import head as oh
from . import refshim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# Full-model cases: BASELINE.json configs at CPU-sized batches (both geometries for resnest26d).
MODEL_CASES = {
    "cfg1_mnist_resnet18_260": dict(args=dict(model="resnet18", dataset="MNIST", channel=512, num_classes=10,
                                              slots_per_class=1, power=1, to_k_layer=1, loss_status=1, lambda_value=1.0),
                                    batch=4, cin=1, size=260),
    "cfg2_resnest26d_pos_260": dict(args=dict(model="resnest26d", dataset="ImageNet", channel=2048, num_classes=10,
                                              slots_per_class=1, power=2, to_k_layer=3, loss_status=1, lambda_value=1.0),
                                    batch=2, cin=3, size=260),
    "cfg2_resnest26d_pos_224": dict(args=dict(model="resnest26d", dataset="ImageNet", channel=2048, num_classes=10,
                                              slots_per_class=1, power=2, to_k_layer=3, loss_status=1, lambda_value=1.0),
                                    batch=3, cin=3, size=224),
    "cfg3_resnest26d_neg_224": dict(args=dict(model="resnest26d", dataset="ImageNet", channel=2048, num_classes=10,
                                              slots_per_class=1, power=2, to_k_layer=3, loss_status=-1, lambda_value=1.0),
                                    batch=2, cin=3, size=224),
    "cfg4_context30_224": dict(args=dict(model="resnest26d", dataset="ConText", channel=2048, num_classes=30,
                                         slots_per_class=1, power=2, to_k_layer=3, loss_status=1, lambda_value=0.2),
                               batch=2, cin=3, size=224),
    "cfg5_cub200x2_224": dict(args=dict(model="resnest26d", dataset="CUB200", channel=2048, num_classes=200,
                                        slots_per_class=2, power=2, to_k_layer=3, loss_status=1, lambda_value=1.0),
                              batch=2, cin=3, size=224),
    # the README's resnest50d runs (README.md:189-228), dumped from the reference's own resnest50d
    "f4_resnest50d_224": dict(args=dict(model="resnest50d", dataset="ImageNet", channel=2048, num_classes=10,
                                        slots_per_class=1, power=2, to_k_layer=3, loss_status=1, lambda_value=1.0),
                              batch=2, cin=3, size=224),
}

# Head-only cases (SlotAttention.forward on synthetic (B,n,64) features): shapes / edge cases.
HEAD_CASES = {
    "head_s10_n81_l3": dict(C=10, spc=1, L=3, n=81, B=8, ls=1, power=2),
    "head_s10_n49_l1_neg": dict(C=10, spc=1, L=1, n=49, B=8, ls=-1, power=1),
    "head_s30_n81_l3": dict(C=30, spc=1, L=3, n=81, B=5, ls=1, power=2),
    "head_s90_spc3_n81": dict(C=30, spc=3, L=3, n=81, B=3, ls=1, power=2),
    "head_s400_spc2_n49": dict(C=200, spc=2, L=3, n=49, B=3, ls=1, power=2),
    "head_s1_n1_b1": dict(C=1, spc=1, L=2, n=1, B=1, ls=1, power=3),
    "head_s7_n64_b1": dict(C=7, spc=1, L=3, n=64, B=1, ls=-1, power=2),
}


def head_inputs(case, seed=7):
    r = np.random.RandomState(seed)
    x = np.maximum(r.standard_normal((case["B"], case["n"], 64)), 0).astype(np.float32) * 0.5
    fs = int(round(case["n"] ** 0.5))
    pe = oh.sine_pe(64, fs, fs).reshape(64, -1).t().contiguous() if fs * fs == case["n"] else \
        torch.from_numpy(r.uniform(-1, 1, (case["n"], 64)).astype(np.float32))
    x = torch.from_numpy(x)
    return x + pe, x


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run_model_case(name, case, check):
    torch.manual_seed(0)
    ref = refshim.reference_slot_model(**case["args"])
    sd = fill_state_dict(ref.state_dict(), seed=0)
    ref.load_state_dict(sd, strict=True)
    x = synth_images(case["batch"], case["cin"], case["size"], case["size"])
    tgt = synth_labels(case["batch"], case["args"]["num_classes"])
    fs = {260: 9, 224: 7}[case["size"]]
    ref.feature_size = fs                                    # the reference hard-wires 9
    feats = {}
    hk = ref.backbone.register_forward_hook(lambda m, i, o: feats.__setitem__("f", o.detach().clone()))
    hs = ref.slot.register_forward_hook(lambda m, i, o: feats.__setitem__("slot", (o[0].detach().clone(), o[1].detach().clone())))
    with torch.no_grad(), refshim.capture_sigmoid() as cap:
        out, (loss, nll, attn_loss) = ref(x, tgt)
    hk.remove(); hs.remove()
    S = case["args"]["num_classes"] * case["args"]["slots_per_class"]
    attn = [o for o in cap.outs if o.dim() == 3 and o.shape[1] == S][-1]
    feat = feats["f"].view(case["batch"], case["args"]["channel"], fs, fs)
    logits = feats["slot"][0]

    ref64 = refshim.reference_slot_model(**case["args"]).double()
    ref64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()})
    ref64.feature_size = fs
    hs64 = ref64.slot.register_forward_hook(lambda m, i, o: feats.__setitem__("slot64", o[0].detach().clone()))
    with torch.no_grad(), refshim.capture_sigmoid() as cap64:
        out64 = ref64(x.double())
    hs64.remove()
    attn64 = [o for o in cap64.outs if o.dim() == 3 and o.shape[1] == S][-1]      # the reference's own fp32-vs-fp64 attention floor
    logits64 = feats["slot64"]

    a = case["args"]
    o = ob.slot_model_forward(a["model"], sd, x, num_classes=a["num_classes"], slots_per_class=a["slots_per_class"],
                              loss_status=a["loss_status"], power=a["power"], lambda_value=a["lambda_value"], target=tgt,
                              return_attn=True)
    ofeat = ob.backbone_features(a["model"], sd, x)
    diffs = dict(feat=rel_err(ofeat, feat), logits=rel_err(o["logits"], logits), log_probs=rel_err(o["log_probs"], out),
                 attn=float((o["attn"] - attn).abs().max()), loss=abs(float(o["loss"] - loss)),
                 noise_floor_log_probs=rel_err(out, out64), noise_floor_attn=float((attn.double() - attn64).abs().max()))
    print(f"[{name}] oracle-vs-reference {json.dumps(diffs)}")
    if check:
        assert diffs["feat"] < 1e-5 and diffs["logits"] < 1e-4 and diffs["attn"] < 1e-3, diffs
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, name + ".npz"),
        meta=json.dumps(dict(kind="model", args=a, batch=case["batch"], cin=case["cin"], size=case["size"], fs=fs,
                             oracle_vs_reference=diffs)),
        log_probs=out.numpy(), log_probs64=out64.numpy(), logits=logits.numpy(), logits64=logits64.numpy(), attn=attn.numpy(),
        attn64=attn64.numpy(),
        losses=np.array([float(loss), float(nll), float(attn_loss)], dtype=np.float32),
        feat_sample=feat[:, ::64].numpy(), feat_abs_mean=np.float32(feat.abs().mean()),
        target=tgt.numpy())


def run_head_case(name, case, check):
    torch.manual_seed(0)
    ref = refshim.reference_slot_attention(case["C"], case["spc"], 64, loss_status=case["ls"], power=case["power"],
                                           to_k_layer=case["L"])
    sd = fill_state_dict(ref.state_dict(), seed=3)
    ref.load_state_dict(sd, strict=True)
    x_pe, x = head_inputs(case)
    with torch.no_grad(), refshim.capture_sigmoid() as cap:
        logits, loss = ref(x_pe, x)
    S = case["C"] * case["spc"]
    attn = [o for o in cap.outs if o.dim() == 3 and o.shape[1] == S][-1]
    ref64 = refshim.reference_slot_attention(case["C"], case["spc"], 64, loss_status=case["ls"], power=case["power"],
                                             to_k_layer=case["L"]).double()
    ref64.load_state_dict({k: v.double() for k, v in sd.items()})
    with torch.no_grad(), refshim.capture_sigmoid() as cap64:
        logits64, loss64 = ref64(x_pe.double(), x.double())
    attn64 = [o for o in cap64.outs if o.dim() == 3 and o.shape[1] == S][-1]
    ol, oloss, oattn = oh.xslot_forward(sd, x_pe, x, num_classes=case["C"], slots_per_class=case["spc"],
                                        loss_status=case["ls"], power=case["power"], return_attn=True)
    diffs = dict(logits=rel_err(ol, logits), attn=float((oattn - attn).abs().max()), loss=abs(float(oloss - loss)),
                 noise_floor_logits=rel_err(logits, logits64), noise_floor_attn=float((attn.double() - attn64).abs().max()))
    print(f"[{name}] oracle-vs-reference {json.dumps(diffs)}")
    if check:
        assert diffs["logits"] < max(1e-4, 20 * diffs["noise_floor_logits"]), diffs
    vis = oh.vis_maps_u8(attn, num_classes=case["C"], slots_per_class=case["spc"], vis_id=0) \
        if int(case["n"] ** 0.5) ** 2 == case["n"] and case["n"] > 1 else np.zeros((0,), np.uint8)
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, name + ".npz"),
        meta=json.dumps(dict(kind="head", case=case, oracle_vs_reference=diffs)),
        logits=logits.numpy(), logits64=logits64.numpy(), attn=attn.numpy(), attn64=attn64.numpy().astype(np.float64),
        loss=np.float32(loss), loss64=np.float64(loss64), vis=vis)


def run_pe_case():
    refshim.install_shims()
    from sloter.utils.position_encode import build_position_encoding
    pe = build_position_encoding("sine", 64)
    out = {}
    for h, w in ((9, 9), (7, 7), (8, 8), (5, 11)):
        t = pe(torch.zeros(1, 64, h, w))[0]
        assert float((t - oh.sine_pe(64, h, w)).abs().max()) < 1e-6
        out[f"pe_{h}x{w}"] = t.numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "pe_sine.npz"), **out)
    print("[pe_sine] ok")


def run_preprocess_case():
    """f2: the reference's ToTensor + Normalize (dataset/transform_func.py:52-67,87-106) followed by engine.py:25's cast
    on seeded uint8 images -> tests/golden/preprocess_u8.npz (PIL Resize excluded: images arrive at the final size)."""
    refshim.install_shims()
    import types
    for mod in ("imgaug", "imgaug.augmenters", "matplotlib", "matplotlib.pyplot"):           # harness-side stub: tools/image_aug.py:1 imports the
        sys.modules.setdefault(mod, types.ModuleType(mod))   # (uninstalled) augmentation library; 'val' never uses it
    sys.modules["imgaug"].augmenters = sys.modules["imgaug.augmenters"]
    from dataset import transform_func as tf
    out = {}
    for ds, c in (("ImageNet", 3), ("MNIST", 1)):
        val = tf.make_transform(argparse.Namespace(dataset=ds, img_size=24, aug=False), "val")
        normalize = val.transforms[-1]                       # Compose([ToTensor(), Normalize(...)])
        img = np.random.RandomState(11 + c).randint(0, 256, size=(3, 24, 20, c)).astype(np.uint8)
        img[0, 0, 0], img[0, 0, 1] = 0, 255
        ref = torch.stack([normalize(im if c > 1 else im[:, :, 0]) for im in img]).to(torch.float32)
        mean, std = {"ImageNet": ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]),
                     "MNIST": ([0.1307], [0.3081])}[ds]
        assert torch.equal(ref, oh.preprocess_u8(img, mean, std)), ds
        out[ds + "_u8"], out[ds + "_f32"] = img, ref.numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "preprocess_u8.npz"), **out)
    print("[preprocess_u8] oracle == reference bit-for-bit")


def run_keys_case():
    """state_dict key/shape contract of the reference modules -> tests/golden/state_dict_keys.json."""
    out = {}
    for name in ("cfg1_mnist_resnet18_260", "cfg2_resnest26d_pos_224", "cfg5_cub200x2_224"):
        ref = refshim.reference_slot_model(**MODEL_CASES[name]["args"])
        out[name] = {k: list(v.shape) for k, v in ref.state_dict().items()}
    ref = refshim.reference_slot_model(model="resnest26d", use_slot=False, num_classes=10)
    out["no_slot_resnest26d"] = {k: list(v.shape) for k, v in ref.state_dict().items()}
    with open(os.path.join(GOLDEN_DIR, "state_dict_keys.json"), "w") as f:
        json.dump(out, f)
    print("[state_dict_keys] ok", {k: len(v) for k, v in out.items()})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    if not refshim.reference_available():
        sys.exit("make_golden needs /root/reference (build container only)")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if not a.only or a.only == "pe":
        run_pe_case()
    if not a.only or a.only == "keys":
        run_keys_case()
    if not a.only or a.only == "preprocess":
        run_preprocess_case()
    for name, case in HEAD_CASES.items():
        if not a.only or a.only in name:
            run_head_case(name, case, a.check)
    for name, case in MODEL_CASES.items():
        if not a.only or a.only in name:
            run_model_case(name, case, a.check)


if __name__ == "__main__":
    main()
