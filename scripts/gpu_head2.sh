#!/bin/bash
# scratch: head-only iteration (tests that reach scouter_head_forward, then timings and per-phase clocks)
timeout 600 python -m pytest tests/test_gpu_head_fused.py tests/test_gpu_conv.py tests/test_gpu_parity.py -m gpu -q -x -k "head or slot_model or small_and_odd or full_size or other_hot or xslot or fused" 2>&1 | tail -3
timeout 120 python scripts/bench_head.py --fs 7 --prof 2>&1 | grep "per-CTA" | cut -c1-1100
timeout 120 python scripts/bench_head.py --fs 7 2>&1 | tail -1
timeout 120 python scripts/bench_head.py --fs 9 2>&1 | tail -1
