#!/bin/bash
# per-kernel-family device time of one forward under a few environment variants
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$label.csv python scripts/profile_forward.py --math tc > gpurun_out/prof_$label.log 2>&1
  python - "$label" <<'PY'
import csv, sys
label=sys.argv[1]
rows=[r for r in csv.reader(open(f'gpurun_out/launches_{label}.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
tot=0; agg={}
for r in rows[1:]:
    v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    tot+=v
    k=r[ki].split('(')[0][-26:]
    agg[k]=agg.get(k,0)+v
print(label, 'total', round(tot), {k:round(v) for k,v in agg.items() if v>100})
PY
}
run base FOO=1
run chunk1000 SCOUTER_UMMA_CHUNK=1000
run bn64 SCOUTER_UMMA_BN=64
run nohalo SCOUTER_NO_HALO=1
