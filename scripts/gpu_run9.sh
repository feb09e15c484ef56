#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "push or k_of_p or timeout" > gpurun_out/pytest_gpu9.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu9.log | tail -5 | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/pytest_multi.log | tail -8 | cut -c1-200
N=2
WIRES=bf16,fp8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29506 scripts/bw_sweep.py > gpurun_out/bw_sweep_n${N}.log 2>&1
grep -E "^\{" gpurun_out/bw_sweep_n${N}.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['bytes'], 'ours_bf16 %.1f us (%.0f GB/s in)  fp8 %.1f us  nccl %.1f us  x%.2f' % (r.get('ours_bf16_us', -1), r.get('ours_bf16_ps_ingress_GBs', -1), r.get('ours_fp8_us', -1), r['nccl_us'], r.get('speedup_vs_nccl_bf16') or 0))
"
grep -E "Error|error" gpurun_out/bw_sweep_n${N}.log | head -5
