set -u
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_kernels.py -q 2>&1 | tail -8
timeout 600 python scripts/bench_conv2.py > gpurun_out/r2c11_bench_conv2.log 2>&1; tail -12 gpurun_out/r2c11_bench_conv2.log
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-baseline 2> gpurun_out/r2c11_$name.err | tail -1 > gpurun_out/r2c11_$name.json
  python -c "import json; d=json.load(open('gpurun_out/r2c11_$name.json')); print('$name', round(d['ms_per_step'],4), 'ms/step', d['final_loss'], d['e2e']['value'] if d['e2e'] else None, d['exposed_comm_ms'])" || tail -3 gpurun_out/r2c11_$name.err; }
b default A=1
b default2 A=1
timeout 300 python bench.py --steps 40 --warmup 5 2> gpurun_out/r2c11_full.err | tail -1 > gpurun_out/r2c11_full.json; python -c "
import json; d=json.load(open('gpurun_out/r2c11_full.json')); print('full', d['ms_per_step'], d['nccl_baseline'], d['vs_nccl_baseline'])"
timeout 300 python bench.py --network VGG16 --steps 20 --warmup 5 --no-baseline 2> gpurun_out/r2c11_vgg.err | tail -1 > gpurun_out/r2c11_vgg.json; python -c "
import json; d=json.load(open('gpurun_out/r2c11_vgg.json')); print('vgg16', d['ms_per_step'], d['value'], d['first_loss'], d['final_loss'])" || tail -5 gpurun_out/r2c11_vgg.err
timeout 300 python bench.py --network ResNet50ImageNet --dataset ImageNet --batch-size 128 --grad-dtype fp8 --steps 10 --warmup 3 --no-baseline 2> gpurun_out/r2c11_r50.err | tail -1 > gpurun_out/r2c11_r50.json; python -c "
import json; d=json.load(open('gpurun_out/r2c11_r50.json')); print('r50 imagenet fp8', d['ms_per_step'], d['value'], d['first_loss'], d['final_loss'])" || tail -5 gpurun_out/r2c11_r50.err
