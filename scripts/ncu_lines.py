"""Aggregate an ncu SASS source page by CUDA source line (debug tool).
python scripts/ncu_lines.py gpurun_out/ncu_head_fused.ncu-rep scouter_b200/libscouter_b200.so head_fused_kernel
Maps the i-th SASS instruction of the kernel to the line nvdisasm reports for it (same cubin => same order)."""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os

rep, lib, kern = sys.argv[1:4]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
data = rows[hdr_i + 1:]
isrc, iss, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
lines = []
for f in os.listdir(tmp):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kern not in txt:
        continue
    cur, infn = 0, False
    for ln in txt.splitlines():
        if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
            infn = kern in ln
        if not infn:
            continue
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
print("sass rows", len(data), "nvdisasm instrs", len(lines))
agg = {}
for i, r in enumerate(data):
    key = lines[i] if i < len(lines) else ("?", 0)
    a = agg.setdefault(key, [0, 0])
    a[0] += int(r[iss] or 0)
    a[1] += int(r[iex] or 0)
tot = sum(v[0] for v in agg.values())
src = {}
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    if f not in src:
        for root, _, files in os.walk("scouter_b200/csrc"):
            if f in files:
                src[f] = open(os.path.join(root, f)).read().splitlines()
    text = src.get(f, [""] * (l + 1))[l - 1].strip()[:90] if f in src and l > 0 else ""
    print(f"{v[0]:6d} {100.0 * v[0] / tot:5.1f}%  inst {v[1]:9d}  {f}:{l}  {text}")
