#!/bin/bash
# accumulation-chunk sweep on one box: speed and the parity figures that move with it
for i in 1 2; do
for c in 2 3 4; do
  export SCOUTER_UMMA_CHUNK=$c
  r=$(python bench.py --no-eager --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.0f img/s backbone %.3f ms' % (d['value'], d['roofline_backbone']['ms']))")
  echo "chunk $c run $i: $r"
done
done
for c in 2 3 4; do
  export SCOUTER_UMMA_CHUNK=$c
  echo "chunk $c:"; python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "reference_golden and not slot_attention" 2>&1 | grep -E "math=1" | sed -E 's/.*(cfg[0-9a-z_]+|f4[0-9a-z_]+) math=1: logits err ([0-9.e-]+).*log-probs err ([0-9.e-]+), attn err ([0-9.e-]+).*/   \1 lp \3 attn \4/'
done
