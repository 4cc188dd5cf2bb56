#!/bin/bash
# Round 2: memcheck over the reworked conv kernels (fp16 main product, folded-tap halo with resident weights), the new aux kernels
# (through one model parity case) and smoke(); racecheck over the folded halo cases.
mkdir -p gpurun_out
F='^  File\|Host Frame\|^=========     by\|^=========         '
{
echo "== memcheck: tests/test_gpu_conv.py (MATH_TC cases)"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "1-B or activation_scale" 2>&1 | grep -v "$F" | tail -6
echo "== memcheck: one model parity case (cfg 2 @224, both math modes) + smoke()"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cfg2_resnest26d_pos_224 and reference_golden" 2>&1 | grep -v "$F" | tail -6
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "$F" | tail -4
echo "== racecheck: folded-tap halo cases"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "1-B1_17x23 or 1-B2_20x20 or 1-B2_56x56_c64-128_k3g2" 2>&1 | grep -v "$F" | tail -6
} > gpurun_out/r02_sanitizer_conv.txt 2>&1
cat gpurun_out/r02_sanitizer_conv.txt
