#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "relu or resnet18_step or conv_autograd" > gpurun_out/pytest_gpu3.log 2>&1; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu3.log | tail -8
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --backend sm100 > gpurun_out/bench_sm100.log 2>&1; tail -1 gpurun_out/bench_sm100.log | cut -c1-330
PSB200_TC_CONV=0 timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --backend sm100 > gpurun_out/bench_sm100_noconv.log 2>&1; tail -1 gpurun_out/bench_sm100_noconv.log | cut -c1-330
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 800 --csv --log-file gpurun_out/launches_sm100.csv python bench.py --gpus 1 --steps 2 --warmup 3 --backend sm100 --no-graph --skip-e2e > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
