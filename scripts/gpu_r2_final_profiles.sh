#!/bin/bash
# Round-2 evidence refresh after the conv rework: parity record, precision-policy table, accumulator-bias probe.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv.py -q -m gpu -s 2>&1 | grep -E "math=|presplit|activation scale|backbone features|passed|failed" > gpurun_out/r02_parity_gpu.txt; tail -1 gpurun_out/r02_parity_gpu.txt
timeout 900 python scripts/precision_policy.py > gpurun_out/r02_precision_policy.txt 2>&1; grep -c img/s gpurun_out/r02_precision_policy.txt
{ python scripts/probes/accum_bias_probe.py; SCOUTER_UMMA_CHUNK=4 python scripts/probes/accum_bias_probe.py; SCOUTER_UMMA_CHUNK=1 python scripts/probes/accum_bias_probe.py; } > gpurun_out/r02_accum_bias_probe.txt 2>&1; tail -3 gpurun_out/r02_accum_bias_probe.txt
