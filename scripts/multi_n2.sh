# 2-GPU validation: every multigpu test that fits on 2 devices (NVLS in-switch reduce, resume, overlap+graph, async pull, ...) + A/B of the NVLS reduce
set -u
export PYTHONPATH=.
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -12
for nvls in 0 1; do
  PSB200_NVLS_REDUCE=$nvls timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29540+nvls)) \
    bench.py --gpus 2 --steps 30 --warmup 5 --topology sharded --skip-e2e --no-baseline 2> gpurun_out/n2_sharded_nvls$nvls.err | tail -1 > gpurun_out/n2_sharded_nvls$nvls.json
  python -c "import json; d=json.load(open('gpurun_out/n2_sharded_nvls$nvls.json')); print('sharded nvls=$nvls', d['ms_per_step'], 'ms/step', d['replica_check'], d['exposed_comm_ms'], d['first_loss'], d['final_loss'])" || tail -5 gpurun_out/n2_sharded_nvls$nvls.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29550 bench.py --gpus 2 --steps 30 --warmup 5 2> gpurun_out/n2_default.err | tail -1 > gpurun_out/n2_default.json
python -c "import json; d=json.load(open('gpurun_out/n2_default.json')); print('ps n=2', d['ms_per_step'], 'ms/step', d['value'], d['replica_check'], d['exposed_comm_ms'], d['first_loss'], d['final_loss'], d['nccl_baseline'])" || tail -5 gpurun_out/n2_default.err
