#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_tc.csv python scripts/profile_forward.py --math tc > gpurun_out/prof_tc.log 2>&1
tail -3 gpurun_out/prof_tc.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_tc.csv')) if len(r)>5]
hdr=rows[0]; 
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
tot=0; 
for r in rows[1:]:
    v=float(r[vi].replace(',','')); 
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    tot+=v
    print(f"{v:10.1f} us  {r[ki][:90]}")
print('total us', tot)
PY
