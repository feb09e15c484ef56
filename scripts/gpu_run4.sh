#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_gpu4.log | tail -30 | cut -c1-200
timeout 600 python scripts/bench_conv.py > gpurun_out/bench_conv.log 2>&1; cat gpurun_out/bench_conv.log | cut -c1-600
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --backend sm100 > gpurun_out/bench_sm100.log 2>&1; tail -1 gpurun_out/bench_sm100.log | cut -c1-330
PSB200_TC_CONV=0 timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --backend sm100 > gpurun_out/bench_sm100_noconv.log 2>&1; tail -1 gpurun_out/bench_sm100_noconv.log | cut -c1-330
