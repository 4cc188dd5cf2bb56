#!/bin/bash
# memcheck over the new kernels (fused head geometries + the TS-mode flat convs); debug tool
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_head_fused.py -m gpu -q -x 2>&1 | grep -v "^  File\|Host Frame\|^=========     by\|^=========         " | tail -12
echo "rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "1-B3_56x56_c64-256_k1g1 or 1-B4_7x7_c512-2048_k1g1 or 1-B2_14x14_c256-320_k1g1 or 1-B300_1x1 or 1-B1_17x17_c64-128" 2>&1 | grep -v "^  File\|Host Frame\|^=========     by\|^=========         " | tail -8
