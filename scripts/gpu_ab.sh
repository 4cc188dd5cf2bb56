#!/bin/bash
# A/B on ONE box: default library vs scouter_b200/libscouter_b200_ab.so (built with an experiment macro), alternating runs.
mkdir -p gpurun_out
for i in 1 2; do
  for v in base ab; do
    if [ $v = ab ]; then export SCOUTER_B200_LIB=$PWD/scouter_b200/libscouter_b200_ab.so; else unset SCOUTER_B200_LIB; fi
    r=$(python bench.py --no-eager --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f img/s backbone %.3f ms clocks %s' % (d['value'], d['roofline_backbone']['ms'], d['clocks']['sm_mhz']))")
    echo "$v run $i: $r"
  done
done
unset SCOUTER_B200_LIB
