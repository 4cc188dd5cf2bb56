"""Times scouter_head_forward alone (debug tool; the judged number comes from bench.py's `roofline`).

python scripts/bench_head.py [--batch 256] [--fs 7] [--classes 10] [--spc 1] [--prof]
--prof loads the -DSCOUTER_PROF build (scripts/prof_roles.py --build-only) and prints the per-phase clocks of the
fused kernel.  SCOUTER_NO_FUSED_HEAD=1 selects the two-kernel path."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scouter_b200 as sb  # noqa: E402
from scouter_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--fs", type=int, default=7)
ap.add_argument("--ch", type=int, default=2048)
ap.add_argument("--classes", type=int, default=10)
ap.add_argument("--spc", type=int, default=1)
ap.add_argument("--layers", type=int, default=3)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--prof", action="store_true")
a = ap.parse_args()
if a.prof:
    L.LIB_PATH = os.path.join(ROOT, "scouter_b200", "libscouter_b200_prof.so")
lib = L.lib()
dev = torch.device("cuda", 0)
B, n, ch = a.batch, a.fs * a.fs, a.ch
r = np.random.RandomState(5)
feat = torch.from_numpy(np.maximum(r.standard_normal((B, n, ch)), 0).astype(np.float32)).to(dev)
w = torch.from_numpy((r.standard_normal((64, ch)) / np.sqrt(ch)).astype(np.float32)).to(dev)
b = torch.from_numpy(0.1 * r.standard_normal(64).astype(np.float32)).to(dev)
m = sb.SlotAttention(a.classes, a.spc, 64, to_k_layer=a.layers, power=2).to(dev).eval()
desc, packed = m.desc_and_pack(dev)
pe = sb.build_position_encoding("sine", 64).table(a.fs, a.fs, dev)
S = a.classes * a.spc
logits = torch.empty(B, a.classes, device=dev)
attn = torch.empty(B, S, n, device=dev)
asum = torch.empty(B, device=dev)
io = L.HeadIO()
io.batch, io.h, io.w, io.channel, io.layout, io.math = B, a.fs, a.fs, ch, L.LAYOUT_NHWC, L.MATH_TC
io.feat, io.conv_w, io.conv_b, io.pe = feat.data_ptr(), w.data_ptr(), b.data_ptr(), pe.data_ptr()
io.logits, io.attn, io.attn_sum, io.x_out = logits.data_ptr(), attn.data_ptr(), asum.data_ptr(), 0
from scouter_b200.plan import split_weights_bf16  # noqa: E402
w_split = split_weights_bf16(w)
io.conv_w_split = w_split.data_ptr()
nbytes = lib.scouter_head_workspace_bytes(C.byref(desc), C.byref(io))
ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
off = (-ws.data_ptr()) % 1024
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run():
    L.check(lib.scouter_head_forward(C.byref(desc), packed.data_ptr(), C.byref(io), ws.data_ptr() + off, nbytes,
                                     torch.cuda.current_stream().cuda_stream))


for _ in range(3):
    run()
torch.cuda.synchronize()
ref_logits = logits.clone()
ts = []
for _ in range(a.iters):
    flush.zero_()                      # evict the features from L2: the head reads them from HBM in the forward too
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
assert torch.equal(ref_logits, logits) or "SCOUTER_HEAD_BLOCKED" in os.environ, "head is not bit-reproducible"
ts = np.array(ts)
# steady-state figure: N launches back to back over a ring of 4 distinct feature buffers (4 x 103 MB > 126 MB L2: every
# launch reads HBM, the evicted lines are clean, launch gaps overlap) -- one event pair around the lot
feats = [feat] + [feat.clone() for _ in range(3)]
def run_ring(k):
    io.feat = feats[k % 4].data_ptr()
    run()
for k in range(8):
    run_ring(k)
torch.cuda.synchronize()
NB2B = 40
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(NB2B):
    run_ring(k)
e1.record()
torch.cuda.synchronize()
b2b = e0.elapsed_time(e1) * 1e3 / NB2B
io.feat = feat.data_ptr()
alg = B * n * ch * 4 + B * a.classes * 4 + B * S * n * 4
peak = 6545.9
print(f"head B={B} n={n} S={S} L={a.layers} fused={'SCOUTER_NO_FUSED_HEAD' not in os.environ}: median {np.median(ts):.1f} us, "
      f"min {ts.min():.1f} us; algorithmic {alg / 1e6:.1f} MB -> {alg / np.median(ts) / 1e3:.0f} GB/s = "
      f"{alg / np.median(ts) / 1e3 / peak:.3f} of {peak} GB/s | back-to-back x{NB2B} over 4 buffers: {b2b:.1f} us/launch = {alg / b2b / 1e3 / peak:.3f}")
if a.prof:
    units = B
    buf = (C.c_ulonglong * (256 * 32))()
    rc = lib.scouter_prof_read_head(buf, 256 * 32)
    arr = np.array(buf[:], dtype=np.float64).reshape(256, 32)[:min(units, 256)]
    names = {0: "phaseA", 1: "mlp", 2: "loop", 3: "prodA.wait_empty", 4: "prodW.wait_done", 5: "iss0.wait_cempty", 6: "iss0.wait_opfull",
             7: "split0.wait_fullA", 8: "split0.wait_done", 9: "split0.wait_fullW", 10: "split0.tmem_st",
             19: "own.stream_end", 24: "own.ho_regs", 25: "own.ho_xt", 20: "own.handover", 21: "own.to_tmem", 22: "own.mlp_done", 23: "own.keys_done", 12: "end.loop_exit", 13: "end.logit_dots", 14: "end.sync", 15: "end.stores"}
    act = arr[arr[:, 28] > 0]
    print(f"globaltimer: first CTA entry -> last CTA exit {(act[:, 29].max() - act[:, 28].min()) / 1e3:.1f} us; entry skew {(act[:, 28].max() - act[:, 28].min()) / 1e3:.1f} us; "
          f"exit skew {(act[:, 29].max() - act[:, 29].min()) / 1e3:.1f} us; per-CTA entry->exit clocks mean {act[:, 27].mean():.0f} max {act[:, 27].max():.0f}; entry->phaseA {act[:, 26].mean():.0f}")
    print("per-CTA clocks (mean / max):", {names.get(i, i): (int(arr[:, i].mean()), int(arr[:, i].max())) for i in range(26) if arr[:, i].any()})
    tb = (C.c_ulonglong * (128 * 8))()
    lib.scouter_trace_read_head(tb, 128 * 8)
    full = np.array(tb[:], dtype=np.int64).reshape(128, 8)
    print("loop phases per iteration (CTA 0): dots_mma normalise update_mma update_operand gates_mma readout+cell")
    for it in range(3):
        print(it, " ".join(f"{int(v):7d}" for v in full[64 + it][:6]))
    print("loop trace (CTA 0, clocks since loop start): dots[issue commit done] own[ld bar] upd[issue commit done] own[sig] P4 gates[issue commit done] rz cell end")
    for it in range(3):
        row = np.concatenate([full[70 + 2 * it], full[71 + 2 * it]])
        order = [0, 1, 2, 3, 4, 8, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15]
        print(it, " ".join(f"{int(row[k]):6d}" for k in order))
    tr = full[:ch // 32]
    t0 = tr[tr > 0].min()
    print("trace (CTA 0, clocks since first event): kb prodA prodW sp.fullA sp.done sp.opfull is.opfull is.fullW is.commit")
    for kb in range(0, tr.shape[0], 8):
        print(kb, " ".join(f"{int(v - t0):7d}" for v in tr[kb]))
