#!/bin/bash
# Round artifacts: full GPU test-suite, smoke, bench (both arms), launch list, ncu capture of the fused head.  Outputs under gpurun_out/.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "math=|err |passed|failed|FAILED|presplit" | tee gpurun_out/pytest_gpu_full.log | tail -8
echo "== bench"; timeout 600 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/bench_final.json | cut -c1-300
tail -2 gpurun_out/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_ref_final.json | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_final.csv python scripts/profile_forward.py --math tc > gpurun_out/prof_tc.log 2>&1
bash scripts/gpu_ncu_fused.sh
for fs in 7 9; do timeout 120 python scripts/bench_head.py --fs $fs 2>&1 | tail -1; SCOUTER_NO_FUSED_HEAD=1 timeout 120 python scripts/bench_head.py --fs $fs 2>&1 | tail -1; done | tee gpurun_out/bench_head.log
