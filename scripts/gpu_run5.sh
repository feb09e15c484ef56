#!/bin/bash
mkdir -p gpurun_out
python scripts/conv_one.py 1 0 > gpurun_out/halo_mode0.log 2>&1; cat gpurun_out/halo_mode0.log | tail -12
python scripts/conv_one.py 1 1 > gpurun_out/halo_mode1.log 2>&1; cat gpurun_out/halo_mode1.log | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_conv_tap_l1 python scripts/conv_one.py 0 > gpurun_out/ncu_conv1.log 2>&1; tail -2 gpurun_out/ncu_conv1.log
SHAPE=128,16,16,128,3,1,1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_conv_tap_l2 python scripts/conv_one.py 0 > gpurun_out/ncu_conv2.log 2>&1; tail -2 gpurun_out/ncu_conv2.log
ls -la gpurun_out/*.ncu-rep
