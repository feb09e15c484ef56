#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu_full.log | tail -8 | cut -c1-200
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for v in 0 1; do
PSB200_TC_CONV=$v timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 --skip-e2e 2>gpurun_out/err_tc$v.log | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('tc_conv=$v', r['ms_per_step'], r['final_loss'], r['gpu_launches'])" || tail -5 gpurun_out/err_tc$v.log
done
