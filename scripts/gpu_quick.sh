#!/bin/bash
mkdir -p gpurun_out
echo "== conv tests"; timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_conv.log
echo "== parity tests"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/pytest_parity.log
echo "== bench tc"; timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_tc.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'head', d['roofline']['ms'], d['roofline']['frac'], 'bb', d['roofline_backbone']['ms'])"
tail -3 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_tc3x.csv python scripts/profile_forward.py --math tc > gpurun_out/prof_tc.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_tc3x.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
tot=0; agg={}
for r in rows[1:]:
    v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    tot+=v
    k=r[ki].split('(')[0][-30:]
    agg[k]=agg.get(k,0)+v
print({k:round(v) for k,v in agg.items()})
print('last 4:', [(r[ki].split('(')[0][-28:], r[vi]) for r in rows[-4:]])
print('total us', tot)
PY
