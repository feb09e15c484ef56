# 1 GPU, the round-end sequence: whole GPU suite, smoke(), the bench in both arms
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py --steps 30 --warmup 5 2> gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json
python -c "import json; d=json.load(open('gpurun_out/final_bench.json')); print('bench', d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['first_loss'], d['final_loss'], d['exposed_comm_ms'], d['replica_check'], d['vs_nccl_baseline'])" || tail -5 gpurun_out/final_bench.err
