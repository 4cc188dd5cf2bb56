#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_cfg5.py 2>&1 | tail -30 | tee gpurun_out/debug_cfg5.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_tc3x.csv python scripts/profile_forward.py --math tc > gpurun_out/prof_tc.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_tc3x.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
tot=0
for r in rows[1:]:
    v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    tot+=v
    if 'umma' in r[ki]: print(f"{v:10.1f} us  {r[ki][:60]}")
print('total us', tot)
PY
