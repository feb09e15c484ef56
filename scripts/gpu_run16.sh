#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "resnet18_step" 2>&1 | tail -2
for i in 1 2; do
for v in 1 0; do
PSB200_WGRAD_STREAM=$v timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 --skip-e2e 2>gpurun_out/err_$v.log | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('wgrad_stream=$v', r['ms_per_step'], r['final_loss'])" || tail -5 gpurun_out/err_$v.log
done; done
