#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
for v in "" "--no-stage"; do
timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 $v 2>gpurun_out/err.log | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('stage flag [$v]', r['ms_per_step'], 'e2e', r['e2e']['ms_per_step'], r['final_loss'])" || tail -5 gpurun_out/err.log
done; done
