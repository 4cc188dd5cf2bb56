#!/bin/bash
# A/B on one box by environment switch: usage gpu_ab_env.sh VAR  (runs bench.py alternately without / with VAR=1)
for i in 1 2 3; do
  for v in base $1; do
    if [ $v = base ]; then unset $1; else export $1=1; fi
    r=$(python bench.py --no-eager --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f img/s backbone %.3f ms launches %d clocks %s' % (d['value'], d['roofline_backbone']['ms'], d['gpu_launches'], d['clocks']['sm_mhz']))")
    echo "$v run $i: $r"
  done
done
unset $1
