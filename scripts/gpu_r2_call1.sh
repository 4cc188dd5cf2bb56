#!/bin/bash
# Round-2 first GPU call: read-bandwidth probes, unit peaks, first hardware run of the f1 draft kernels, baseline bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/c1_gpu.txt
echo "== hbm read probe"; timeout 300 scripts/probes/hbm_read_probe 2>&1 | tee gpurun_out/hbm_read_probe.txt | tail -70
echo "== unit peaks"; timeout 300 scripts/probes/unit_peaks_probe 2>&1 | tee gpurun_out/unit_peaks.json
echo "== f1 drafts (plain)"; timeout 900 python -m pytest tests -m gpu_draft -x -q 2>&1 | tail -25 | tee gpurun_out/f1_drafts.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_start.json 2> gpurun_out/bench_r2_start.err; tail -c 3000 gpurun_out/bench_r2_start.json; tail -5 gpurun_out/bench_r2_start.err
echo "== head prof"; timeout 200 python scripts/bench_head.py --fs 7 --prof 2>&1 | head -8 | tee gpurun_out/head_prof_start.txt
