#!/bin/bash
# Launch list of one forward (cfg 3, B=256, 224^2) -> per-op table.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/r02_launches_b256.csv python scripts/profile_forward.py --math tc > gpurun_out/prof_fwd.log 2>&1
tail -2 gpurun_out/prof_fwd.log
python scripts/analyze_launches.py gpurun_out/r02_launches_b256.csv > gpurun_out/r02_launch_table_b256.txt 2>&1; tail -3 gpurun_out/r02_launch_table_b256.txt
