set -u
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -8
for v in 16 0; do
  PSB200_TC_HALO_PITCH=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-baseline 2> gpurun_out/r2c9_bench.err | tail -1 > gpurun_out/r2c9_bench_pitch$v.json
  python -c "import json; d=json.load(open('gpurun_out/r2c9_bench_pitch$v.json')); print('halo pitch=$v', d['ms_per_step'], 'ms/step', d['first_loss'], d['final_loss'], d['exposed_comm_ms'], d['gpu_launches'])" || tail -5 gpurun_out/r2c9_bench.err
done
PSB200_TC_CONV=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-baseline 2> gpurun_out/r2c9_bench.err | tail -1 > gpurun_out/r2c9_bench_cudnn.json
python -c "import json; d=json.load(open('gpurun_out/r2c9_bench_cudnn.json')); print('cudnn convs', d['ms_per_step'], 'ms/step')"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c9_launches.csv python bench.py --steps 2 --warmup 3 --no-baseline --skip-e2e --no-graph > gpurun_out/r2c9_ncu_bench.log 2>&1; tail -2 gpurun_out/r2c9_ncu_bench.log | cut -c1-300
