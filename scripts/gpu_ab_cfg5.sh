#!/bin/bash
# A/B on one box for the S=400 head (cfg 5): default library vs the variants scouter_b200/libscouter_b200_ab*.so
for i in 1 2; do
  for v in base $(ls scouter_b200/libscouter_b200_ab*.so 2>/dev/null); do
    if [ $v = base ]; then unset SCOUTER_B200_LIB; else export SCOUTER_B200_LIB=$PWD/$v; fi
    r=$(python bench.py --config cfg5 --no-eager --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f img/s head %.1f us backbone %.3f ms' % (d['value'], 1e3*d['roofline']['ms'], d['roofline_backbone']['ms']))")
    echo "$v run $i: $r"
  done
done
unset SCOUTER_B200_LIB
