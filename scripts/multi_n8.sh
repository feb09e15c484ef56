# 8-GPU runs: headline (reference topology co-located / literal 1 PS + 7 workers), PS pull mode with in-switch reduction, sharded, strong scaling,
# BASELINE configs 3-5 (ResNet-50/ImageNet-shape fp8 wire, VGG-16, bandwidth sweep), the unmodified reference on 8 GPUs
set -u
export PYTHONPATH=.
run() { name=$1; port=$2; shift 2; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 20 --warmup 5 $EXTRA 2> gpurun_out/n8_$name.err | tail -1 > gpurun_out/n8_$name.json
  python -c "import json; d=json.load(open('gpurun_out/n8_$name.json')); print('$name', round(d['ms_per_step'],4), 'ms/step', round(d['value']), d['replica_check'], 'exposed', d['exposed_comm_ms'], d['comm_incl_peer_wait_ms_max'], d['first_loss'], d['final_loss'], (d.get('nccl_baseline') or {}).get('value'), d['clocks']['reasons'])" || tail -5 gpurun_out/n8_$name.err; }
EXTRA="" run default 29750 A=1
EXTRA="--no-baseline --topology sharded" run sharded 29752 A=1
EXTRA="--no-baseline" run pull 29751 PSB200_PS_PULL=1
EXTRA="--no-baseline --network VGG16" run vgg16 29756 A=1
EXTRA="--no-baseline --network ResNet50ImageNet --dataset ImageNet --batch-size 128 --grad-dtype fp8 --steps 10 --warmup 3" run r50_fp8 29757 A=1
EXTRA="--no-baseline --global-batch 1024" run strong1024 29755 A=1
EXTRA="--no-baseline --no-colocate" run nocolo 29753 A=1
EXTRA="--no-baseline --num-aggregate 7" run k7of8 29759 A=1
timeout 200 python -m pytest tests/test_gpu_multi.py -q --tb=short -k "k_of_p" > gpurun_out/n8_pytest_kofp.log 2>&1; tail -4 gpurun_out/n8_pytest_kofp.log
WIRES=bf16 timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29760 scripts/bw_sweep.py > gpurun_out/n8_bw_sweep.log 2>&1; tail -14 gpurun_out/n8_bw_sweep.log | cut -c1-400
