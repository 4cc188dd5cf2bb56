#!/bin/bash
# Bench lines of the BASELINE configs on ONE GPU (cfg 4 / cfg 5 are quoted on 4 / 8 GPUs: see gpu_bench_multi.sh)
mkdir -p gpurun_out
for c in cfg3 cfg2 cfg4 cfg5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_${c}_1gpu.json 2> gpurun_out/bench_${c}_1gpu.err || tail -5 gpurun_out/bench_${c}_1gpu.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${c}_1gpu.json"))
    r = d["roofline"]; e = d.get("gpu_eager_baseline") or {}
    print("${c}", "value %.0f img/s" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "module_call %.0f" % d["module_call"]["value"],
          "| head %.1f us frac %.3f (single flushed %.1f us) launches %d" % (1e3 * r["ms"], r["frac"], 1e3 * r["ms_single_flushed"], r["launches"]),
          "| backbone %.2f ms %.0f TF/s frac %.3f" % (d["roofline_backbone"]["ms"], d["roofline_backbone"]["achieved"], d["roofline_backbone"]["frac"]),
          "| cpu %.1f img/s (%s, %d cores)" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"]),
          "| eager tf32 %.0f fp32 %.0f" % (e.get("default_flags", {}).get("value", 0), e.get("tf32_off", {}).get("value", 0)), d["clocks"])
except Exception as ex:
    print("${c} failed:", ex)
PY
done
