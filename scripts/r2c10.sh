set -u
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_kernels.py -q 2>&1 | tail -8
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-baseline --skip-e2e 2> gpurun_out/r2c10_$name.err | tail -1 > gpurun_out/r2c10_$name.json
  python -c "import json; d=json.load(open('gpurun_out/r2c10_$name.json')); print('$name', round(d['ms_per_step'],4), 'ms/step', d['final_loss'])" || tail -3 gpurun_out/r2c10_$name.err; }
b default A=1
b prio0 PSB200_STEP_PRIORITY=0
b wg120 PSB200_WGRAD_SMS=120
b wg96 PSB200_WGRAD_SMS=96
b wg74 PSB200_WGRAD_SMS=74
b default2 A=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_conv" -s 2 -c 1 -f -o gpurun_out/r2c10_halo2_l1 python scripts/ncu_one.py halo2_l1 2>&1 | tail -1
