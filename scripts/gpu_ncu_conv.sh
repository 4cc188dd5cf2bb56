#!/bin/bash
# ncu --set full of two conv kernels: the layer-4 1x1 (flat, TS mode) and the layer-3 3x3 (halo, BN = 128)
mkdir -p gpurun_out
M="--metrics sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_sectors.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"
timeout 600 ncu --set full --clock-control none -k regex:conv_umma_kernel -s 2 -c 1 -o gpurun_out/r02_ncu_flat_l4 -f python scripts/conv_case.py --hw 7 --cin 1024 --cout 2048 --k 1 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none -k regex:conv3x3_halo_kernel -s 2 -c 1 -o gpurun_out/r02_ncu_halo_l3 -f python scripts/conv_case.py --hw 28 --cin 256 --cout 512 --k 3 --groups 2 2>&1 | tail -1
for f in flat_l4 halo_l3; do ncu -i gpurun_out/r02_ncu_$f.ncu-rep --page raw --csv > gpurun_out/r02_ncu_${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep
