#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -k "litmus or async or ddp" > gpurun_out/pytest_multi4.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/pytest_multi4.log | tail -8 | cut -c1-200
grep -E "^E  " gpurun_out/pytest_multi4.log | grep -v "^E *$" | head -24 | cut -c1-260
N=4
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 30 --warmup 5 "${@:3}" > gpurun_out/$2.log 2>&1; tail -1 gpurun_out/$2.log | cut -c1-330; grep -E "Error|error:" gpurun_out/$2.log | head -3; }
run 29501 bench_n4_default
run 29502 bench_n4_sharded --topology sharded --skip-e2e
WIRES=bf16 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29506 scripts/bw_sweep.py > gpurun_out/bw_sweep_n${N}.log 2>&1
grep -E "^\{" gpurun_out/bw_sweep_n${N}.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['bytes'], 'ours_bf16 %.1f us (%.0f GB/s in)  nccl %.1f us  x%.2f' % (r.get('ours_bf16_us', -1), r.get('ours_bf16_ps_ingress_GBs', -1), r['nccl_us'], r.get('speedup_vs_nccl_bf16') or 0))
"
