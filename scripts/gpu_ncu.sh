#!/bin/bash
mkdir -p gpurun_out
cap() {  # name kernel-regex skip
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
     -k regex:"$2" -s $3 -c 1 -o gpurun_out/ncu_$1 -f python scripts/profile_forward.py --math tc > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
# final build of round 1 (profiles/r01_ncu_summary.md, second table): the two dominant deep-layer kernels
cap halo128_layer3_0 conv3x3_halo_kernel 6
cap flat128_layer4_down conv_umma_kernel 30
# earlier captures
cap halo32_conv1_3 conv3x3_halo_kernel 0
cap halo128_layer2_0 conv3x3_halo_kernel 4
cap flat128_layer1_conv3 conv_umma_kernel 2
cap xslot_loop xslot_loop_kernel 0
cap head_proj conv_umma_kernel 21
ls -la gpurun_out/*.ncu-rep
