"""Stand-alone evaluator: polls ``<model_dir>model_step_<k·eval_freq>``, loads the state_dict into a fresh model,
prints average NLL + Prec@1/Prec@5 (reference ``/root/reference/src/distributed_evaluator.py:58-114``).  It shares
nothing with the training job but the filesystem.  Additions: ``max_evals`` / ``timeout_seconds`` so it can terminate
(the reference loops forever), ``num_classes`` from the dataset (the reference hard-codes 10), GPU evaluation."""
from __future__ import annotations

import os
import time

import torch

from ..models import build_model
from ..utils import checkpoint as ckpt
from ..utils.metrics import accuracy


class DistributedEvaluator(object):
    def __init__(self, **kwargs):
        self._cur_step = 0
        self._model_dir = kwargs['model_dir']
        self._eval_freq = int(kwargs['eval_freq'])
        self._eval_batch_size = kwargs['eval_batch_size']
        self.network_config = kwargs['network']
        self._num_classes = kwargs.get('num_classes', 10)
        self._device = torch.device(kwargs.get('device', 'cpu'))
        self._poll = float(kwargs.get('poll_seconds', 10.0))
        self._max_evals = int(kwargs.get('max_evals', 0))
        self._timeout = float(kwargs.get('timeout_seconds', 0.0))
        self.results = []

    def evaluate(self, validation_loader):
        self._next_step_to_fetch = self._cur_step + self._eval_freq
        last_progress = time.time()
        while True:
            model_dir_ = self._model_dir_generator(self._next_step_to_fetch)
            if os.path.isfile(model_dir_):
                self._load_model(model_dir_)
                print("Evaluator evaluating results on step {}".format(self._next_step_to_fetch))
                self.results.append((self._next_step_to_fetch,) + self._evaluate_model(validation_loader))
                self._next_step_to_fetch += self._eval_freq
                last_progress = time.time()
                if self._max_evals and len(self.results) >= self._max_evals:
                    return self.results
            else:
                if self._timeout and time.time() - last_progress > self._timeout:
                    return self.results
                time.sleep(self._poll)

    @torch.no_grad()
    def _evaluate_model(self, test_loader):
        self.network.eval()
        test_loss, p1, p5, nb, n = 0.0, 0.0, 0.0, 0, 0
        for data, y_batch in test_loader:
            data, target = data.to(self._device).float(), y_batch.to(self._device)
            if self._device.type == "cuda":
                data = data.contiguous(memory_format=torch.channels_last) if data.dim() == 4 else data
            output = self.network(data).float()
            test_loss += torch.nn.functional.nll_loss(torch.log_softmax(output, dim=1), target, reduction='sum').item()
            a1, a5 = accuracy(output, target, topk=(1, 5))
            p1 += a1.item(); p5 += a5.item(); nb += 1; n += target.numel()
        prec1, prec5, test_loss = p1 / max(nb, 1), p5 / max(nb, 1), test_loss / max(n, 1)
        print('Test set: Average loss: {:.4f}, Prec@1: {} Prec@5: {}'.format(test_loss, prec1, prec5))
        return test_loss, prec1, prec5

    def _load_model(self, file_path):
        self.network = build_model(self.network_config, num_classes=self._num_classes)
        ckpt.load_model_into(self.network, file_path)
        self.network.to(self._device)
        if self._device.type == "cuda":
            self.network = self.network.to(memory_format=torch.channels_last)

    def _model_dir_generator(self, next_step_to_fetch):
        return ckpt.model_path(self._model_dir, next_step_to_fetch)
