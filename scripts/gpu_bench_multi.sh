#!/bin/bash
# bench lines of the multi-GPU BASELINE configs: cfg 4 on 4 GPUs, cfg 5 on 8 GPUs (usage: gpu_bench_multi.sh <cfg> <ngpu>)
mkdir -p gpurun_out
c=$1; n=$2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --config $c --gpus $n --steps 10 --warmup 3 \
    > gpurun_out/bench_${c}_${n}gpu.json 2> gpurun_out/bench_${c}_${n}gpu.err || tail -5 gpurun_out/bench_${c}_${n}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${c}_${n}gpu.json"))
print("${c} x${n}: value %.0f img/s, ms/step %.2f, e2e %.0f, head %.1f us, backbone %.2f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"], 1e3 * d["roofline"]["ms"], d["roofline_backbone"]["ms"]), d["clocks"])
PY
