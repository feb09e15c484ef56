#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"reduce_update_kernel|push_kernel" -s 3 -c 2 -o gpurun_out/prof_comm python bench.py --gpus 1 --steps 1 --warmup 3 --no-graph --skip-e2e > gpurun_out/ncu_comm.log 2>&1; tail -1 gpurun_out/ncu_comm.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bn_stats_kernel|bn_apply_kernel|bn_bwd_reduce_kernel|bn_bwd_apply_kernel" -s 128 -c 8 -o gpurun_out/prof_bn_l1 python bench.py --gpus 1 --steps 1 --warmup 3 --no-graph --skip-e2e > gpurun_out/ncu_bn.log 2>&1; tail -1 gpurun_out/ncu_bn.log | cut -c1-200
