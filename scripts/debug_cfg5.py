import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import load_golden
import scouter_b200 as sb
from scouter_b200 import _lib as L
from scouter_b200.synth import fill_state_dict, synth_images
from scouter_b200.synth import make_args
dev = torch.device("cuda", 0)
for name in ("cfg4_context30_224", "cfg5_cub200x2_224", "cfg2_resnest26d_pos_224"):
    z, meta = load_golden(name)
    for math in (0, 1):
        m = sb.SlotModel(make_args(**meta["args"]))
        m.load_state_dict(fill_state_dict(m.state_dict(), seed=0))
        m = m.to(dev).eval(); m.math = math; m.keep_attn = True
        x = synth_images(meta["batch"], meta["cin"], meta["size"], meta["size"]).to(dev)
        with torch.no_grad():
            out = m(x).cpu()
        attn = m.last_attn.cpu()
        ra = torch.from_numpy(z["attn"]); r32 = torch.from_numpy(z["log_probs"]); r64 = torch.from_numpy(z["log_probs64"]).float()
        ea = (attn - ra).abs()
        rows_bad = (ea.max(2).values > 0.05).nonzero().tolist()
        print(f"{name} math={math}: lp err vs ref32 {float((out-r32).abs().max()):.3e} vs ref64 {float((out-r64).abs().max()):.3e} "
              f"(ref32 vs ref64 {float((r32-r64).abs().max()):.3e}); attn max err {float(ea.max()):.3e}; bad rows {rows_bad[:8]} n={len(rows_bad)}")
        for (b, i) in rows_bad[:3]:
            print("    row", b, i, "ours", attn[b, i, :6].tolist(), "ref", ra[b, i, :6].tolist())
