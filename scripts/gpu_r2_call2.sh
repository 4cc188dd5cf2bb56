#!/bin/bash
mkdir -p gpurun_out
echo "== hbm read probe (dirty vs clean L2)"; timeout 300 scripts/probes/hbm_read_probe 2>&1 | tee gpurun_out/hbm_read_probe2.txt | tail -80
echo "== head prof"; timeout 200 python scripts/bench_head.py --fs 7 --prof 2>&1 | head -12 | cut -c1-1500 | tee gpurun_out/head_prof_start.txt
