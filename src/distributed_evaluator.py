"""Evaluator entry — same name/flags as the reference (``/root/reference/src/distributed_evaluator.py``)."""
import argparse

import torch

import _path  # noqa: F401
from ps_pytorch_b200.config import add_eval_args, num_classes_for
from ps_pytorch_b200.parallel.evaluator import DistributedEvaluator
from ps_pytorch_b200.utils.data import prepare_data

if __name__ == "__main__":
    args = add_eval_args(argparse.ArgumentParser(description='Distributed Evaluator'))
    args.batch_size = args.eval_batch_size
    device = torch.device("cuda" if args.enable_gpu and torch.cuda.is_available() else "cpu")
    _, test_loader = prepare_data(args)
    evaluator_nn = DistributedEvaluator(network=args.network, model_dir=args.model_dir, eval_freq=args.eval_freq,
                                        eval_batch_size=args.eval_batch_size, num_classes=num_classes_for(args.dataset),
                                        device=device, poll_seconds=args.poll_seconds, max_evals=args.max_evals,
                                        timeout_seconds=args.timeout_seconds)
    evaluator_nn.evaluate(validation_loader=test_loader)
