#!/bin/bash
for b in 32 64 128 256; do
  timeout 120 python scripts/bench_head.py --fs 7 --batch $b --prof 2>&1 | grep -v "^[0-9]" | tail -3 | head -2 | cut -c1-260
done
echo "== NA sweep (nw 6)"
for na in 2 3 4 6; do SCOUTER_HEAD_NA=$na timeout 120 python scripts/bench_head.py --fs 7 --prof 2>&1 | grep "per-CTA" | cut -c1-120; done
echo "== NA 8 nw 4 / na 8 nw 3"
SCOUTER_HEAD_NW=4 timeout 120 python scripts/bench_head.py --fs 7 --prof 2>&1 | grep "per-CTA" | cut -c1-120
SCOUTER_HEAD_NW=3 timeout 120 python scripts/bench_head.py --fs 7 --prof 2>&1 | grep "per-CTA" | cut -c1-120
