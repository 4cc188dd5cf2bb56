#!/bin/bash
mkdir -p gpurun_out
echo "== conv tests"; timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/pytest_conv.log
echo "== parity tests"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -60 | tee gpurun_out/pytest_parity.log
echo "== bench tc"; timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_tc.json
tail -5 gpurun_out/bench.err
echo "== bench tc_fast"; timeout 600 python bench.py --steps 10 --warmup 3 --math tc_fast 2>gpurun_out/bench.err | tee gpurun_out/bench_tc_fast.json
