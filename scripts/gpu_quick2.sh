#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_tc.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'head', d['roofline']['ms'], d['roofline']['frac'], 'bb', d['roofline_backbone']['ms'], d['clocks'])"
tail -2 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_final.csv python scripts/profile_forward.py --math tc > gpurun_out/prof_tc.log 2>&1
