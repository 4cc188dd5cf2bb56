#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stream or host" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench_tc.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['synchronous_call_value'], 'head', d['roofline']['ms'], d['roofline']['frac'], d['clocks'])"
tail -2 gpurun_out/bench.err
