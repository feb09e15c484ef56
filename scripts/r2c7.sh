set -u
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -8
timeout 600 python scripts/bench_conv2.py > gpurun_out/r2c7_bench_conv2.log 2>&1; tail -12 gpurun_out/r2c7_bench_conv2.log
PSB200_TC_CONV=1 timeout 300 python bench.py --steps 30 --warmup 5 2> gpurun_out/r2c7_bench.err | tail -1 > gpurun_out/r2c7_bench.json
python -c "import json; d=json.load(open('gpurun_out/r2c7_bench.json')); print('all ours', d['ms_per_step'], 'ms/step', d['config']['conv'], d['final_loss'], d['e2e'])" || tail -5 gpurun_out/r2c7_bench.err
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
