#!/bin/bash
# first full GPU pass: kernel numerics, smoke, 1-GPU bench (sm100 / library backends), stock baseline
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --backend sm100 > gpurun_out/bench_sm100.log 2>&1; tail -3 gpurun_out/bench_sm100.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --backend library > gpurun_out/bench_library.log 2>&1; tail -2 gpurun_out/bench_library.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --impl nccl-baseline > gpurun_out/bench_nccl1.log 2>&1; tail -2 gpurun_out/bench_nccl1.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --backend sm100 --grad-dtype fp8 > gpurun_out/bench_sm100_fp8.log 2>&1; tail -2 gpurun_out/bench_sm100_fp8.log
