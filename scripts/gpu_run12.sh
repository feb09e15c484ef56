#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not tc_" > gpurun_out/pytest_gpu12.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu12.log | tail -8 | cut -c1-200
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-330; grep -o '"gpu_launches": [0-9]*' gpurun_out/bench_default.log
PSB200_BN_DETERMINISTIC=1 timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --skip-e2e > gpurun_out/bench_det.log 2>&1; tail -1 gpurun_out/bench_det.log | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 700 --csv --log-file gpurun_out/launches_v2.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-graph --skip-e2e > gpurun_out/ncu_bench.log 2>&1
