#!/usr/bin/env bash
# round 2, GPU call 2: new CTA-pair tap-table kernel (fprop + strided dgrad), per-layer table, all-ours step vs cuDNN step
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv2 or weight_transpose or all_passes_ours or halo_cta_pair or finalize_wide" 2>&1 | tail -15 | tee gpurun_out/r2c2_tests.log
timeout 600 python scripts/bench_conv2.py > gpurun_out/r2c2_bench_conv2.log 2>&1; tail -14 gpurun_out/r2c2_bench_conv2.log
for tc in 1 0; do
  PSB200_TC_CONV=$tc timeout 300 python bench.py --steps 30 --warmup 5 --skip-e2e 2> gpurun_out/r2c2_bench_tc$tc.err | tail -1 > gpurun_out/r2c2_bench_tc$tc.json
  python -c "import json; d=json.load(open('gpurun_out/r2c2_bench_tc$tc.json')); print('PSB200_TC_CONV=$tc', d['ms_per_step'], 'ms/step', d['config']['conv'], d['final_loss'])" || tail -5 gpurun_out/r2c2_bench_tc$tc.err
done
