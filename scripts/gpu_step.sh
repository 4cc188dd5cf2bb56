#!/bin/bash
# one optimisation step on the GPU: conv unit tests, role stall profile, short bench
mkdir -p gpurun_out
echo "== conv tests"; timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_conv.log
echo "== roles"; timeout 300 python scripts/prof_roles.py > gpurun_out/prof_roles.txt 2>&1; grep -E "^==|issuer.total|split.total|epi.total" gpurun_out/prof_roles.txt
echo "== bench tc"; timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_tc.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'head', d['roofline']['ms'], d['roofline']['frac'], 'bb', d['roofline_backbone']['ms'])"
tail -3 gpurun_out/bench.err
if [ "$1" == "full" ]; then
echo "== parity tests"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_parity.log
fi
