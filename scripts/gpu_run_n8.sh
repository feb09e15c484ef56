#!/bin/bash
# 8-GPU pass: scaling point of the headline bench (ps / sharded / fp8 / reference topology), NCCL baseline, bandwidth sweep
mkdir -p gpurun_out
N=${1:-8}
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 30 --warmup 5 "${@:3}" > gpurun_out/$2.log 2>&1; tail -1 gpurun_out/$2.log | cut -c1-600; grep -E "Error|error:" gpurun_out/$2.log | head -3; }
run 29501 bench_n${N}_ps --backend sm100
run 29502 bench_n${N}_sharded --backend sm100 --topology sharded
run 29503 bench_n${N}_ps_fp8 --backend sm100 --grad-dtype fp8 --skip-e2e
run 29504 bench_n${N}_nccl --impl nccl-baseline
run 29505 bench_n${N}_ps_nocolo --backend sm100 --no-colocate --skip-e2e
WIRES=bf16,fp8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29506 scripts/bw_sweep.py > gpurun_out/bw_sweep_n${N}.log 2>&1
grep -E "^\{" gpurun_out/bw_sweep_n${N}.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['bytes'], 'ours_bf16 %.1f us (%.0f GB/s in)  fp8 %.1f us  nccl %.1f us  x%.2f' % (r.get('ours_bf16_us', -1), r.get('ours_bf16_ps_ingress_GBs', -1), r.get('ours_fp8_us', -1), r['nccl_us'], r.get('speedup_vs_nccl_bf16') or 0))
"
grep -E "Error|error" gpurun_out/bw_sweep_n${N}.log | head -5
