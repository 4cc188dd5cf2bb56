#!/bin/bash
# Round-2 closing run: smoke, full GPU suite, default bench (both arms), launch table.  Outputs under gpurun_out/.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee gpurun_out/r02_pytest_gpu_final.txt
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/r02_bench_reference.json | cut -c1-400; tail -1 gpurun_out/bench_ref.err
echo "== bench"; timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r02_bench_final.json | cut -c1-300; tail -1 gpurun_out/bench.err
bash scripts/gpu_launch_table.sh | tail -1
