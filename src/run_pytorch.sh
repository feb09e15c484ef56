#!/bin/bash
# Launch line of the reference (src/run_pytorch.sh: mpirun -n 3 ... ResNet18 / Cifar10 / batch 1024 / lr 0.1 / momentum 0.9 /
# compress) on this framework's launcher: one process per GPU, rank 0 = parameter server.
#   NPROC=8 bash run_pytorch.sh            # 8 GPUs of one B200 box (default: all visible GPUs, or 3 CPU ranks without GPUs)
#   overrides: NETWORK DATASET BATCH MAX_STEPS EVAL_FREQ COLOCATE TRAIN_DIR
cd "$(dirname "$0")"
NGPU=$(python -c "import torch; print(torch.cuda.device_count())" 2>/dev/null || echo 0)
NPROC=${NPROC:-$([ "$NGPU" -gt 0 ] && echo "$NGPU" || echo 3)}
GPU_FLAG=$([ "$NGPU" -gt 0 ] && echo "--enable-gpu=true" || echo "--enable-gpu=")
python -m ps_pytorch_b200.launch -n "$NPROC" distributed_nn.py \
--lr=0.1 \
--momentum=0.9 \
--max-steps=${MAX_STEPS:-100000} \
--epochs=100 \
--network=${NETWORK:-ResNet18} \
--dataset=${DATASET:-Cifar10} \
--batch-size=${BATCH:-1024} \
--comm-type=Bcast \
--num-aggregate=$((NPROC-1)) \
--mode=normal \
--eval-freq=${EVAL_FREQ:-20} \
--compress-grad=compress \
--colocate-worker=${COLOCATE:-false} \
$GPU_FLAG \
--train-dir=${TRAIN_DIR:-output/models/}
