#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/pytest_multi.log | tail -12
grep -E "^E  " gpurun_out/pytest_multi.log | grep -v "^E   *$" | head -40 | cut -c1-250
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 30 --warmup 5 "${@:3}" > gpurun_out/$2.log 2>&1; tail -1 gpurun_out/$2.log | cut -c1-900; grep -E "Error|error" gpurun_out/$2.log | head -3; }
run 29501 bench_n2_ps --backend sm100
run 29502 bench_n2_sharded --backend sm100 --topology sharded
run 29503 bench_n2_ps_fp8 --backend sm100 --grad-dtype fp8
run 29505 bench_n2_ps_nocolo --backend sm100 --no-colocate
