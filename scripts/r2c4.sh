set -u
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv or bn or basic_block or resnet18" 2>&1 | tail -6
timeout 600 python scripts/bench_conv2.py > gpurun_out/r2c4_bench_conv2.log 2>&1; tail -12 gpurun_out/r2c4_bench_conv2.log
for v in "1 1" "1 0" "0 1"; do set -- $v
  PSB200_TC_CONV=$1 PSB200_CONV_BN_STATS=$2 timeout 300 python bench.py --steps 30 --warmup 5 --skip-e2e 2> gpurun_out/r2c4_bench.err | tail -1 > gpurun_out/r2c4_bench_tc$1_st$2.json
  python -c "import json; d=json.load(open('gpurun_out/r2c4_bench_tc$1_st$2.json')); print('tc=$1 stats=$2', d['ms_per_step'], 'ms/step', d['config']['conv'], d['final_loss'])" || tail -5 gpurun_out/r2c4_bench.err
done
