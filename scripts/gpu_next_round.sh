#!/usr/bin/env bash
# First GPU call of the next round: validate the three gated kernels written after round 1's GPU budget ran out, and A/B them.
#   gpurun --timeout 900 -- 'bash scripts/gpu_next_round.sh 1'        # 1 GPU: wide BN finalize, CTA-pair halo conv
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_next_round.sh 2' # 2 GPUs: NVLS in-switch reduction (sharded topology)
# Everything lands in gpurun_out/next_round_*.  Each step runs under `timeout` so a protocol bug cannot hang the box.
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  PSB200_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "finalize_wide or halo_cta_pair" 2>&1 | tail -15 | tee gpurun_out/next_round_tests_1gpu.log
  for mode in default wide; do
    env $([ "$mode" = wide ] && echo PSB200_BN_FINALIZE=wide) timeout 200 python bench.py --steps 40 --warmup 5 --skip-e2e 2>&1 | tail -1 > gpurun_out/next_round_bench_finalize_$mode.json
    python -c "import json; d=json.load(open('gpurun_out/next_round_bench_finalize_$mode.json')); print('$mode', d['ms_per_step'], 'ms/step', d['clocks'])"
  done
  # all-tc convolutions (halo kernel where eligible) with and without the CTA-pair variant; the per-tap-only number of round 1 was 9.13 ms
  for pair in 0 1; do
    PSB200_TC_CONV=1 PSB200_TC_CONV_ALGO=1 PSB200_TC_HALO_PAIR=$pair timeout 240 python bench.py --steps 30 --warmup 5 --skip-e2e 2>&1 | tail -1 > gpurun_out/next_round_bench_tcconv_pair$pair.json
    python -c "import json; d=json.load(open('gpurun_out/next_round_bench_tcconv_pair$pair.json')); print('tc conv, halo pair=$pair', d['ms_per_step'], 'ms/step')"
  done
else
  PSB200_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_multi.py -q -x -k "nvls" 2>&1 | tail -15 | tee gpurun_out/next_round_tests_${N}gpu.log
  for nvls in 0 1; do
    PSB200_NVLS_REDUCE=$nvls timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+nvls)) \
      bench.py --gpus $N --steps 40 --warmup 5 --topology sharded --skip-e2e 2>&1 | tail -1 > gpurun_out/next_round_bench_sharded_nvls$nvls.json
    python -c "import json; d=json.load(open('gpurun_out/next_round_bench_sharded_nvls$nvls.json')); print('sharded, nvls=$nvls', d['ms_per_step'], 'ms/step')"
  done
fi
