#!/bin/bash
# FIRST GPU run of the row-f1 draft kernels (scouter_b200/csrc/draft/): plain, then under memcheck and racecheck.
# Nothing here is part of `pytest -m gpu`; see tests/test_gpu_draft_kernels.py.
mkdir -p gpurun_out
echo "== gpu_draft"; timeout 900 python -m pytest tests -m gpu_draft -x -q 2>&1 | tail -15 | tee gpurun_out/f1_drafts.log
echo "== memcheck"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu_draft -x -q -k each_kernel_group 2>&1 \
  | grep -v "^  File\|Host Frame\|^=========     by\|^=========         " | tail -12 | tee gpurun_out/f1_drafts_memcheck.log
echo "== racecheck"; timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu_draft -x -q -k "training_program and resnet18" 2>&1 \
  | grep -v "^  File\|Host Frame\|^=========     by\|^=========         " | tail -12 | tee gpurun_out/f1_drafts_racecheck.log
