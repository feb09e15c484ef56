set -u
export PYTHONPATH=.
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "all_passes_ours or tc_conv_fprop or halo_cta_pair" 2>&1 | tail -5
timeout 600 python scripts/bench_conv2.py > gpurun_out/r2c3_bench_conv2.log 2>&1; tail -12 gpurun_out/r2c3_bench_conv2.log
PSB200_TC_CONV=1 timeout 300 python bench.py --steps 30 --warmup 5 --skip-e2e 2> gpurun_out/r2c3_bench.err | tail -1 > gpurun_out/r2c3_bench_tc1.json
python -c "import json; d=json.load(open('gpurun_out/r2c3_bench_tc1.json')); print('all ours', d['ms_per_step'], 'ms/step', d['config']['conv'], d['final_loss'])" || tail -5 gpurun_out/r2c3_bench.err
for k in halo2_l1 dgrad_l2s2 wgrad_l1 wgrad_l4 conv2_l3 halo2_l2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_conv|tc_wgrad" -s 2 -c 1 -f -o gpurun_out/r2c3_$k python scripts/ncu_one.py $k 2>&1 | tail -2
done
