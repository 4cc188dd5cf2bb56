#!/bin/bash
# Fused-head iteration: parity tests that go through scouter_head_forward, then head-only timings.
mkdir -p gpurun_out
if [ "$1" != "notest" ]; then
echo "== head tests"
timeout 900 python -m pytest tests/test_gpu_head_fused.py tests/test_gpu_conv.py tests/test_gpu_parity.py -m gpu -q -s -k "head or slot_model or small_and_odd or full_size or other_hot" 2>&1 | grep -E "fused head|log-prob|err|assert|passed|failed|FAILED" | cut -c1-220 | tail -40 | tee gpurun_out/pytest_head.log
fi
echo "== head bench"
for fs in 7 9; do
  timeout 120 python scripts/bench_head.py --fs $fs 2>&1 | tail -1
done
timeout 120 python scripts/bench_head.py --fs 7 --classes 30 2>&1 | tail -1
timeout 120 python scripts/bench_head.py --fs 7 --prof 2>&1 | sed -n 1,12p | cut -c1-700
