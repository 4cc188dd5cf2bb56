#!/bin/bash
# Launch list of one training step (forward + backward, resnest26d, B=32, 224^2) -> per-kernel totals.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv \
    --log-file gpurun_out/r02_launches_train_b32.csv python scripts/profile_train_step.py > gpurun_out/prof_train.log 2>&1
tail -1 gpurun_out/prof_train.log
python scripts/analyze_launches.py gpurun_out/r02_launches_train_b32.csv > gpurun_out/r02_launch_table_train_b32.txt 2>&1
head -25 gpurun_out/r02_launch_table_train_b32.txt
