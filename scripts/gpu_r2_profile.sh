#!/bin/bash
# Round-2 profile set: launch list of one forward (cfg 3, B=256, 224^2), ncu --set full of the head kernel, at both geometries.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/r02_launches_b256.csv python scripts/profile_forward.py --math tc > gpurun_out/prof_fwd.log 2>&1
tail -2 gpurun_out/prof_fwd.log
python scripts/analyze_launches.py gpurun_out/r02_launches_b256.csv > gpurun_out/r02_launch_table_b256.txt 2>&1; tail -3 gpurun_out/r02_launch_table_b256.txt
for fs in 7 9; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_fused_kernel -s 3 -c 1 -o gpurun_out/r02_ncu_head_fs$fs -f \
      python scripts/bench_head.py --fs $fs --iters 3 > gpurun_out/r02_ncu_head_fs$fs.log 2>&1
  tail -1 gpurun_out/r02_ncu_head_fs$fs.log
  ncu -i gpurun_out/r02_ncu_head_fs$fs.ncu-rep --page raw --csv > gpurun_out/r02_ncu_head_fs${fs}_raw.csv 2>/dev/null
done
cp scouter_b200/libscouter_b200.so gpurun_out/lib_profiled.so
for fs in 7 9; do timeout 120 python scripts/bench_head.py --fs $fs | tail -1; done
timeout 120 python scripts/bench_head.py --fs 7 --classes 30 | tail -1
