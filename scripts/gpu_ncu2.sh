#!/bin/bash
mkdir -p gpurun_out
cap() {  # name kernel-regex skip
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
     -k regex:"$2" -s $3 -c 1 -o gpurun_out/ncu_$1 -f python scripts/profile_forward.py --math tc > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap halo128_layer3_0 conv3x3_halo_kernel 6
cap flat128_layer4_down conv_umma_kernel 30
