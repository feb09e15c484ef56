# 2-GPU re-validation of the three fixes (bound-gradient accumulate vs side-stream wgrad, BN-statistics send over raw peer pointers)
set -u
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "resume_on_nvlink or overlap_with_graph or sharded_and_ddp" 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "bound_gradients" 2>&1 | tail -3
run() { name=$1; port=$2; shift 2; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 30 --warmup 5 $EXTRA 2> gpurun_out/n2_$name.err | tail -1 > gpurun_out/n2_$name.json
  python -c "import json; d=json.load(open('gpurun_out/n2_$name.json')); print('$name', round(d['ms_per_step'],4), 'ms/step', round(d['value']), d['replica_check'], 'exposed', d['exposed_comm_ms'], d['comm_incl_peer_wait_ms_max'], d['first_loss'], d['final_loss'])" || tail -5 gpurun_out/n2_$name.err; }
EXTRA="--no-baseline --overlap" run overlap 29741 A=1
EXTRA="--no-baseline --overlap --no-graph" run overlap_nograph 29742 A=1
