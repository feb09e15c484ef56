#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi.log 2>&1; grep -E "passed|failed|FAILED" gpurun_out/pytest_multi.log | tail -8 | cut -c1-200
grep -E "^E  " gpurun_out/pytest_multi.log | head -20 | cut -c1-250
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 30 --warmup 5 "${@:3}" > gpurun_out/$2.log 2>&1; tail -1 gpurun_out/$2.log | cut -c1-330; grep -E "Error|error:" gpurun_out/$2.log | head -3; }
run 29501 bench_n2_default
run 29502 bench_n2_overlap --overlap --skip-e2e
run 29503 bench_n2_sharded --topology sharded --skip-e2e
