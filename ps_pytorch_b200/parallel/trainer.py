"""Single-machine trainer (the scalability baseline of the reference, ``/root/reference/src/nn_ops.py:29-106`` +
``single_machine.py``): zero_grad → forward → loss → backward → optimizer step → Prec@1/5, validate per epoch.
Here it runs on the same compute core and fused optimizer kernel as the distributed roles, on GPU when asked."""
from __future__ import annotations

import time

import torch

from .arena import FlatBuffer
from .compute import WorkerCompute, resolve_dtype
from ..models import build_model
from ..optim import build_optimizer
from ..utils.metrics import accuracy


class NN_Trainer(object):
    def __init__(self, **kwargs):
        self.batch_size = kwargs['batch_size']
        self.lr = kwargs['learning_rate']
        self.max_epochs = kwargs['max_epochs']
        self.momentum = kwargs['momentum']
        self.network_config = kwargs['network']
        self._device = torch.device(kwargs.get('device', 'cpu'))
        self._dtype = resolve_dtype(kwargs.get('dtype', 'auto'), self._device)
        self._num_classes = kwargs.get('num_classes', 10)
        self._max_steps = kwargs.get('max_steps', 0)
        self._optimizer_name = kwargs.get('optimizer', 'sgd')
        self._log_interval = max(1, kwargs.get('log_interval', 1))
        self._use_graph = kwargs.get('cuda_graph', False)
        self.history = []

    def build_model(self):
        net = build_model(self.network_config, self._num_classes)
        self.compute = WorkerCompute(net, self._device, self._dtype, use_graph=self._use_graph)
        self.network = self.compute.network
        # fp32 master + fused step(grads) — the PS data path, in-process
        self.master = FlatBuffer(self.compute.layout, torch.float32, self._device)
        self.master.flat.copy_(self.compute.weights.flat)
        self.optimizer = build_optimizer(self._optimizer_name, [torch.nn.Parameter(v) for v in self.master.views], self.lr, self.momentum)
        self.optimizer.attach_flat(self.master)
        self.criterion = self.compute.criterion
        return self

    def train_step(self, x, y):
        loss, logits = self.compute.train_step(x, y)
        lowp = None if self.compute.weights.dtype == torch.float32 else self.compute.weights.flat
        self.optimizer.step(grads=self.compute.grads.flat, out_lowp=lowp)
        if lowp is None:
            self.compute.weights.flat.copy_(self.master.flat)
        return loss, logits

    def train_and_validate(self, train_loader, test_loader):
        step = 0
        for i in range(self.max_epochs):
            for batch_idx, (data, y_batch) in enumerate(train_loader):
                t0 = time.time()
                loss, logits = self.train_step(data, y_batch)
                step += 1
                if step % self._log_interval == 0:
                    prec1, prec5 = accuracy(logits.float(), y_batch.to(logits.device), topk=(1, 5))
                    lv = float(loss)
                    self.history.append(lv)
                    print('Train Epoch: {} [{}/{} ({:.0f}%)]\tLoss: {:.6f}  Prec@1: {}  Prec@5: {}  Time Cost: {}'.format(
                        i, batch_idx * len(data), len(train_loader.dataset), 100. * batch_idx / max(1, len(train_loader)), lv,
                        prec1.item(), prec5.item(), time.time() - t0))
                if self._max_steps and step >= self._max_steps:
                    self.validate(test_loader)
                    return
            self.validate(test_loader)

    def validate(self, test_loader):
        loss, p1, p5 = self.compute.evaluate(test_loader)
        print('Test set: Average loss: {:.4f}, Prec@1: {} Prec@5: {}'.format(loss, p1, p5))
        return loss, p1, p5
