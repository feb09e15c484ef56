# 1 GPU: programmatic dependent launch (csrc/pdl.cuh) — the whole single-GPU suite with PDL on, then the step A/B
set -u
export PYTHONPATH=.
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for v in 1 0; do
  PSB200_PDL=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-baseline 2> gpurun_out/r2c13_pdl$v.err | tail -1 > gpurun_out/r2c13_pdl$v.json
  python -c "import json; d=json.load(open('gpurun_out/r2c13_pdl$v.json')); print('PDL=$v', d['ms_per_step'], d['value'], d['e2e']['value'], d['first_loss'], d['final_loss'], d['weight_abs_sum'] if 'weight_abs_sum' in d else '')" || tail -5 gpurun_out/r2c13_pdl$v.err
done
PSB200_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-baseline --network VGG16 2> gpurun_out/r2c13_vgg.err | tail -1 > gpurun_out/r2c13_vgg.json
python -c "import json; d=json.load(open('gpurun_out/r2c13_vgg.json')); print('VGG16 PDL=1', d['ms_per_step'], d['value'], d['first_loss'], d['final_loss'])" || tail -5 gpurun_out/r2c13_vgg.err
