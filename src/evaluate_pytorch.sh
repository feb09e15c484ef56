#!/bin/bash
# Evaluator launch line of the reference (src/evaluate_pytorch.sh): polls --model-dir for model_step_<k*eval_freq>.
cd "$(dirname "$0")"
python distributed_evaluator.py \
--eval-batch-size=${EVAL_BATCH:-10000} \
--eval-freq=${EVAL_FREQ:-20} \
--network=${NETWORK:-ResNet18} \
--dataset=${DATASET:-Cifar10} \
--model-dir=${TRAIN_DIR:-output/models/} \
"$@"
