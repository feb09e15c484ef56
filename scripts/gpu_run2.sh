#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -40
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --backend sm100 > gpurun_out/bench_sm100.log 2>&1; tail -2 gpurun_out/bench_sm100.log | cut -c1-400
PSB200_TC_CONV=0 timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --backend sm100 > gpurun_out/bench_sm100_noconv.log 2>&1; tail -1 gpurun_out/bench_sm100_noconv.log | cut -c1-300
PSB200_TC_CONV=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 900 --csv --log-file gpurun_out/launches_sm100.csv python bench.py --gpus 1 --steps 2 --warmup 3 --backend sm100 --no-graph --skip-e2e > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
