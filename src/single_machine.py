"""Single-machine trainer entry — same name/flags as the reference (``/root/reference/src/single_machine.py``)."""
import argparse

import torch

import _path  # noqa: F401
from ps_pytorch_b200.config import add_single_machine_args, num_classes_for
from ps_pytorch_b200.parallel.trainer import NN_Trainer
from ps_pytorch_b200.parallel.runtime import seed_everything
from ps_pytorch_b200.utils.data import prepare_data

if __name__ == "__main__":
    args = add_single_machine_args(argparse.ArgumentParser(description='Single Machine Trainer'))
    use_cuda = (args.enable_gpu if args.enable_gpu is not None else torch.cuda.is_available()) and not args.no_cuda
    device = torch.device("cuda" if use_cuda and torch.cuda.is_available() else "cpu")
    seed_everything(args.seed)
    if args.backend != "auto":
        from ps_pytorch_b200 import nn as pnn
        pnn.set_backend(args.backend)
    train_loader, test_loader = prepare_data(args, pin=device.type == "cuda", channels_last=device.type == "cuda")
    kwargs = {'batch_size': args.batch_size, 'learning_rate': args.lr, 'max_epochs': args.epochs, 'momentum': args.momentum,
              'network': args.network, 'device': device, 'dtype': args.dtype, 'num_classes': num_classes_for(args.dataset),
              'max_steps': args.max_steps, 'optimizer': args.optimizer, 'log_interval': args.log_interval}
    nn_learner = NN_Trainer(**kwargs)
    nn_learner.build_model()
    nn_learner.train_and_validate(train_loader=train_loader, test_loader=test_loader)
