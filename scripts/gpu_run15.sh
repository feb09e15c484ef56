#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
for v in 1 0; do
PSB200_BN_MASKX=$v timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 --skip-e2e 2>/dev/null | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('maskx=$v', r['ms_per_step'], r['clocks'])"
done; done
