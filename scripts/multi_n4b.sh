# 4 GPUs: the K < P paths (staggered litmus, kill / timeout CLI runs) + the K = 3 of 4 bench
set -u
export PYTHONPATH=.
timeout 700 python -m pytest tests/test_gpu_multi.py -q --tb=long -k "k_of_p or straggler" > gpurun_out/n4b_pytest.log 2>&1; tail -5 gpurun_out/n4b_pytest.log
env timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 4 --steps 30 --warmup 5 --no-baseline --num-aggregate 3 2> gpurun_out/n4_k3of4.err | tail -1 > gpurun_out/n4_k3of4.json
python -c "import json; d=json.load(open('gpurun_out/n4_k3of4.json')); print('k3of4', round(d['ms_per_step'],4), 'ms/step', round(d['value']), d['replica_check'], 'exposed', d['exposed_comm_ms'], d['first_loss'], d['final_loss'])" || tail -30 gpurun_out/n4_k3of4.err
