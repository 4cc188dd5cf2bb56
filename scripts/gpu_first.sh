#!/bin/bash
# First GPU pass: smoke, parity tests, a short bench, kernel launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
echo "== bench reference" ; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_ref.json
