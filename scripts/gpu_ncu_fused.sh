#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_fused_kernel -s 3 -c 1 -o gpurun_out/ncu_head_fused -f python scripts/bench_head.py --fs 7 --iters 3 > gpurun_out/ncu_head_fused.log 2>&1
tail -2 gpurun_out/ncu_head_fused.log
cp scouter_b200/libscouter_b200.so gpurun_out/lib_profiled.so
