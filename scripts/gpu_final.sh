#!/bin/bash
# Round artifacts: full GPU test-suite, smoke, bench (both arms), launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "math=|err |passed|failed|FAILED|presplit" | tee gpurun_out/pytest_gpu_full.log | tail -25
echo "== bench"; timeout 600 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/bench_final.json | cut -c1-400
tail -2 gpurun_out/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_ref_final.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_final.csv python scripts/profile_forward.py --math tc > gpurun_out/prof_tc.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_final.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
tot=0; agg={}
for r in rows[1:]:
    v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    tot+=v
    k=r[ki].split('(')[0][-30:]
    agg[k]=agg.get(k,0)+v
print({k:round(v) for k,v in agg.items()}); print('total us', tot)
PY
