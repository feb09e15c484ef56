# 2-GPU validation (run with: gpurun --gpus 2 -- 'bash scripts/multi_n2.sh'): the whole multi-GPU test file, then the headline bench in its variants and the unmodified reference
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=short > gpurun_out/n2_pytest.log 2>&1; tail -12 gpurun_out/n2_pytest.log
run() { name=$1; port=$2; shift 2; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 30 --warmup 5 $EXTRA 2> gpurun_out/n2_$name.err | tail -1 > gpurun_out/n2_$name.json
  python -c "import json; d=json.load(open('gpurun_out/n2_$name.json')); print('$name', round(d['ms_per_step'],4), 'ms/step', round(d['value']), d['replica_check'], 'exposed', d['exposed_comm_ms'], d['comm_incl_peer_wait_ms_max'], d['first_loss'], d['final_loss'], (d.get('nccl_baseline') or {}).get('value'))" || tail -5 gpurun_out/n2_$name.err; }
EXTRA="" run default 29740 A=1
EXTRA="--no-baseline" run pull 29741 PSB200_PS_PULL=1
EXTRA="--no-baseline --topology sharded" run sharded 29742 A=1
EXTRA="--no-baseline --topology sharded" run sharded_nonvls 29743 PSB200_NVLS_REDUCE=0
EXTRA="--no-baseline --overlap" run overlap 29744 A=1
EXTRA="--no-baseline --no-colocate" run nocolo 29745 A=1
timeout 600 python bench.py --impl reference --gpus 2 --steps 5 --warmup 3 2> gpurun_out/n2_reference.err | tail -1 > gpurun_out/n2_reference.json; cut -c1-400 gpurun_out/n2_reference.json
timeout 600 python scripts/overlap_noise.py 32 2>&1 | tail -12
