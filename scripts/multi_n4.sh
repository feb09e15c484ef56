# 4-GPU validation (gpurun --gpus 4): the tests that need >= 3 GPUs (11 M-element litmus, K < P with rotating stragglers, kill / timeout CLI runs), the 4-GPU headline and K = 3 of 4
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=long -k "(litmus and resnet18) or k_of_p or straggler or overlap_with_graph" > gpurun_out/n4_pytest.log 2>&1; tail -6 gpurun_out/n4_pytest.log
run() { name=$1; port=$2; shift 2; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 4 --steps 30 --warmup 5 $EXTRA 2> gpurun_out/n4_$name.err | tail -1 > gpurun_out/n4_$name.json
  python -c "import json; d=json.load(open('gpurun_out/n4_$name.json')); print('$name', round(d['ms_per_step'],4), 'ms/step', round(d['value']), d['replica_check'], 'exposed', d['exposed_comm_ms'], d['comm_incl_peer_wait_ms_max'], d['first_loss'], d['final_loss'], (d.get('nccl_baseline') or {}).get('value'))" || tail -5 gpurun_out/n4_$name.err; }
EXTRA="" run default 29650 A=1
EXTRA="--no-baseline --num-aggregate 3" run k3of4 29652 A=1
