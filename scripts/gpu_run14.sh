#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "batchnorm or resnet18_step" > gpurun_out/pytest_gpu14.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu14.log | tail -8 | cut -c1-200
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-330
