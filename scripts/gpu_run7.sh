#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_halo -s 2 -c 1 -o gpurun_out/prof_conv_halo_l1 python scripts/conv_one.py 1 > gpurun_out/ncu_conv1.log 2>&1; tail -1 gpurun_out/ncu_conv1.log
SHAPE=128,16,16,128,3,1,1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_halo -s 2 -c 1 -o gpurun_out/prof_conv_halo_l2 python scripts/conv_one.py 1 > gpurun_out/ncu_conv2.log 2>&1; tail -1 gpurun_out/ncu_conv2.log
SHAPE=256,8,8,256,3,1,1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_conv_tap_l3 python scripts/conv_one.py 0 > gpurun_out/ncu_conv3.log 2>&1; tail -1 gpurun_out/ncu_conv3.log
