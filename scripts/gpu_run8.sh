#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_halo -s 2 -c 1 -o gpurun_out/prof_conv_halo_res_l1 python scripts/conv_one.py 1 > gpurun_out/ncu_conv1.log 2>&1; tail -1 gpurun_out/ncu_conv1.log
