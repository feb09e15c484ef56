#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/convergence.py > gpurun_out/convergence.log 2>&1; tail -3 gpurun_out/convergence.log | cut -c1-200
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-1200
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --network VGG16 --skip-e2e > gpurun_out/bench_vgg16.log 2>&1; tail -1 gpurun_out/bench_vgg16.log | cut -c1-330
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --network VGG16 --backend library --skip-e2e > gpurun_out/bench_vgg16_lib.log 2>&1; tail -1 gpurun_out/bench_vgg16_lib.log | cut -c1-330
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --network ResNet50ImageNet --dataset ImageNet --batch-size 128 --grad-dtype fp8 --skip-e2e > gpurun_out/bench_r50.log 2>&1; tail -1 gpurun_out/bench_r50.log | cut -c1-330
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --network ResNet50ImageNet --dataset ImageNet --batch-size 128 --backend library --skip-e2e > gpurun_out/bench_r50_lib.log 2>&1; tail -1 gpurun_out/bench_r50_lib.log | cut -c1-330
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"reduce_update_kernel|push_kernel|bn_stats_kernel|bn_apply_kernel|bn_bwd_reduce_kernel|bn_bwd_apply_kernel" -s 40 -c 12 -o gpurun_out/prof_comm_bn python bench.py --gpus 1 --steps 1 --warmup 3 --no-graph --skip-e2e > gpurun_out/ncu_comm.log 2>&1; tail -1 gpurun_out/ncu_comm.log | cut -c1-200
