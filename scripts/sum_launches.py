"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv):  python scripts/sum_launches.py list.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1.0)
    n = r[kn].split("(")[0][:80]
    agg[n][0] += 1
    agg[n][1] += v
    tot += v
print(f"total {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches (ncu: serialised, cold caches)")
for n, (c, v) in sorted(agg.items(), key=lambda t: -t[1][1])[:top]:
    print(f"{v / 1e3:9.2f} ms  x{c:4d}  {n}")
