# 1 GPU: gradient sinks + fused weights-ready wait + bound-gradient fix; bench; compute-sanitizer over one full step
set -u
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "bound_gradients or without_a_gather or weights_ready_wait or tc_conv_wgrad or stem_tensorcore or weight_transpose or all_passes_ours" 2>&1 | tail -8
timeout 300 python bench.py --steps 30 --warmup 5 2> gpurun_out/r2c12_bench.err | tail -1 > gpurun_out/r2c12_bench.json
python -c "import json; d=json.load(open('gpurun_out/r2c12_bench.json')); print('bench', d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['first_loss'], d['final_loss'], d['exposed_comm_ms'])" || tail -5 gpurun_out/r2c12_bench.err
PSB200_GRAD_SINKS=0 PSB200_FUSE_WAIT=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-baseline 2> gpurun_out/r2c12_bench_off.err | tail -1 > gpurun_out/r2c12_bench_off.json
python -c "import json; d=json.load(open('gpurun_out/r2c12_bench_off.json')); print('bench sinks/fuse off', d['ms_per_step'], d['value'], d['gpu_launches'])" || tail -5 gpurun_out/r2c12_bench_off.err
TOOLS="synccheck racecheck" SAN_TIMEOUT=300 bash scripts/sanitize.sh
