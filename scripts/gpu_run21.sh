#!/bin/bash
mkdir -p gpurun_out
for v in 0 auto 1; do
PSB200_TC_CONV=$v timeout 400 python bench.py --gpus 1 --steps 12 --warmup 3 --network ResNet50ImageNet --dataset ImageNet --batch-size 128 --skip-e2e 2>gpurun_out/err_r50_$v.log | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('r50 tc_conv=$v', r['ms_per_step'], r['value'], r['final_loss'])" || tail -3 gpurun_out/err_r50_$v.log
done
for v in 0 1; do
PSB200_TC_CONV=$v timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 --skip-e2e 2>gpurun_out/err_tc$v.log | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('r18 tc_conv=$v', r['ms_per_step'], r['final_loss'])" || tail -5 gpurun_out/err_tc$v.log
done
