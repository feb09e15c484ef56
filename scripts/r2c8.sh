set -u
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -8
for v in 1 0; do
  PSB200_TC_WGRAD_STREAM=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-baseline 2> gpurun_out/r2c8_bench.err | tail -1 > gpurun_out/r2c8_bench_side$v.json
  python -c "import json; d=json.load(open('gpurun_out/r2c8_bench_side$v.json')); print('wgrad side stream=$v', d['ms_per_step'], 'ms/step', d['config']['conv'], d['first_loss'], d['final_loss'], d['replica_check'], d['exposed_comm_ms'])" || tail -5 gpurun_out/r2c8_bench.err
done
timeout 300 python bench.py --steps 30 --warmup 5 2> gpurun_out/r2c8_bench.err | tail -1 > gpurun_out/r2c8_bench_full.json; cat gpurun_out/r2c8_bench_full.json | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c8_launches.csv python bench.py --steps 2 --warmup 3 --no-baseline --skip-e2e --no-graph > gpurun_out/r2c8_ncu_bench.log 2>&1; tail -2 gpurun_out/r2c8_ncu_bench.log
