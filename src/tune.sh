#!/bin/bash
# Learning-rate sweep of the reference (src/tune.sh): 7 learning rates x ${MAX_TUNING_STEP} steps, then average the
# last-step loss over workers with tiny_tuning_parser.py.  Ranks are supervised by ps_pytorch_b200.launch, so no killall.
cd "$(dirname "$0")"
tune_dir=${TUNE_DIR:-./tune/}
max_tuning_step=${MAX_TUNING_STEP:-100}
nproc=${NPROC:-17}
lrs=${LRS:-"0.0078125 0.015625 0.03125 0.0625 0.125 0.25 0.5"}
mkdir -p "${tune_dir}"
echo "Start parameter tuning ..."
for lr in ${lrs}
do
  echo "Trial running for learning rate: ${lr}"
  python -m ps_pytorch_b200.launch -n ${nproc} distributed_nn.py \
  --lr=${lr} --momentum=0.9 --network=${NETWORK:-ResNet18} --dataset=${DATASET:-Cifar10} --batch-size=${BATCH:-8} --test-batch-size=200 \
  --comm-type=Bcast --num-aggregate=$((nproc-1)) --eval-freq=200 --epochs=10 --max-steps=${max_tuning_step} --log-interval=1 \
  --enable-gpu=${ENABLE_GPU:-} --train-dir=${tune_dir} > ${tune_dir}lr_${lr} 2>&1
  grep "Step: ${max_tuning_step}," ${tune_dir}lr_${lr} > ${tune_dir}lr_${lr}_processing
done
for lr in ${lrs}
do
  echo "Logging out tunning results"
  python tiny_tuning_parser.py --tuning-dir=${tune_dir}lr_${lr}_processing --tuning-lr=${lr} --num-workers=$((nproc-1))
done
