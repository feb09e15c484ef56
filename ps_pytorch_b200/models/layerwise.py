"""Layer-wise backward with in-flight gradient pushes and straggler kill.

Capability parity with the reference's ``LeNetSplit`` / ``ResNetSplit``
(``/root/reference/src/model_ops/lenet.py:39-257``, ``resnet_split.py:142-749``): gradients leave the
worker while the rest of backward is still running, last layer first; between layers the worker
polls the PS kill signal (reference: MPI tag 77, ``resnet_split.py:513-523``) or its own timeout and
abandons the step.

Mechanism (new): the network is cut into *stages* (stem, each residual block, head).  ``forward``
re-leafs the activation at every stage boundary; the backward walks the stages in reverse with
``torch.autograd.backward`` per stage and hands the finished stage's parameter indices to a
*communicator* — any object with ``push_params(indices, step)`` (the NVLink push kernel on a side
stream, or the gloo sender) and ``poll_kill(step) -> bool`` (an abort word in symmetric memory).
"""
from __future__ import annotations

import time
from typing import List, Sequence

import torch
import torch.nn as tnn


class StragglerKilled(RuntimeError):
    """Raised/returned when a worker abandons a step on the PS's kill signal or its own timeout."""


class _NullCommunicator:
    def push_params(self, indices, step):
        pass

    def poll_kill(self, step):
        return False


def default_stages(model: tnn.Module) -> List[List[tnn.Module]]:
    """Execution-ordered stage list.  Each stage is a list of modules applied by ``_run_stage``."""
    from .lenet import LeNet
    from .resnet import ResNet
    if isinstance(model, LeNet):
        return [["conv1"], ["conv2"], ["fc1"], ["fc2"]]
    if isinstance(model, ResNet):
        st = [["stem"]]
        for ln in ("layer1", "layer2", "layer3", "layer4"):
            for i in range(len(getattr(model, ln))):
                st.append([f"{ln}.{i}"])
        st.append(["head"])
        return st
    raise TypeError("no stage map for %s" % type(model).__name__)


class LayerwiseBackwardMixin:
    kill_threshold: float = 7.0

    # ---- stage execution ---------------------------------------------------------------------
    def _run_stage(self, tag: str, x):
        from .lenet import LeNet
        if isinstance(self, LeNet):
            if tag in ("conv1", "conv2"):
                return self.pool(getattr(self, tag)(x), relu=True)
            if tag == "fc1":
                return self.fc1(x.reshape(x.size(0), -1))
            return self.fc2(x)
        if tag == "stem":
            out = self.bn1(self.conv1(x, bn=self.bn1), relu=True)
            return self.stem_pool(out) if self.imagenet else out
        if tag == "head":
            from .resnet import _global_mean
            out = _global_mean(x) if self.imagenet else self.pool(x).reshape(x.size(0), -1)
            return self.linear(out)
        ln, i = tag.split(".")
        return getattr(self, ln)[int(i)](x)

    def _stage_modules(self, tag: str) -> Sequence[tnn.Module]:
        from .lenet import LeNet
        if isinstance(self, LeNet):
            return [getattr(self, tag)]
        if tag == "stem":
            return [self.conv1, self.bn1]
        if tag == "head":
            return [self.linear]
        ln, i = tag.split(".")
        return [getattr(self, ln)[int(i)]]

    def _build_stage_index(self):
        pid = {id(p): i for i, p in enumerate(self.parameters())}
        self._stage_tags = [s[0] for s in default_stages(self)]
        self._stage_param_idx = []
        for tag in self._stage_tags:
            idx = []
            for m in self._stage_modules(tag):
                idx += [pid[id(p)] for p in m.parameters()]
            self._stage_param_idx.append(sorted(idx))

    def forward(self, x):
        if not hasattr(self, "_stage_tags"):
            self._build_stage_index()
        if not torch.is_grad_enabled():
            for tag in self._stage_tags:
                x = self._run_stage(tag, x)
            return x
        # record (input, output) per stage like the reference's self.input / self.output lists
        self.input, self.output = [], []
        for tag in self._stage_tags:
            x = x.detach().requires_grad_(True) if self.input else x
            self.input.append(x)
            x = self._run_stage(tag, x)
            self.output.append(x)
        return x

    # ---- backward variants ---------------------------------------------------------------------
    def _walk(self, g, communicator, cur_step, kill_mode, on_stage=None):
        comm = communicator if communicator is not None else _NullCommunicator()
        t0 = time.time()
        for i in range(len(self.output) - 1, -1, -1):
            if kill_mode == "signal" and comm.poll_kill(cur_step):
                return True
            if kill_mode == "timeout" and (time.time() - t0) > float(self.kill_threshold):
                return True
            out = self.output[i]
            grad = g if i == len(self.output) - 1 else self.input[i + 1].grad
            torch.autograd.backward(out, grad)
            comm.push_params(self._stage_param_idx[i], cur_step)
            if on_stage is not None:
                on_stage(i)
        return False

    def backward_single(self, g):
        """Plain stage-by-stage backward, no communication (reference ``resnet_split.py:730-749``)."""
        self._walk(g, None, 0, None)

    def backward_normal(self, g, communicator=None, req_send_check=None, cur_step=0, compress_grad='compress'):
        """Backward with per-stage pushes overlapped with the remaining backward
        (reference ``resnet_split.py:365-501``).  ``req_send_check``/``compress_grad`` are accepted for
        signature parity; ordering/flow control lives in the communicator."""
        self._walk(g, communicator, cur_step, None)
        return req_send_check

    def backward(self, g, communicator=None, req_send_check=None, cur_step=0):
        """Step-token variant (reference ``resnet_split.py:262-363``): every push carries ``cur_step``
        (here always: the arrival flag *is* the step id, so stale pushes are rejected by value)."""
        self._walk(g, communicator, cur_step, None)
        return req_send_check

    def backward_signal_kill(self, g, communicator=None, req_send_check=None, cur_step=0):
        """Poll the PS kill signal before every stage; returns ``(req_send_check, killed)``
        (reference ``resnet_split.py:503-615``)."""
        killed = self._walk(g, communicator, cur_step, "signal")
        return req_send_check, killed

    def backward_timeout_kill(self, g, communicator=None, req_send_check=None, cur_step=0):
        """Self-abort when the backward has run longer than ``kill_threshold`` seconds
        (reference ``resnet_split.py:617-728``; a stub for LeNet there, implemented here)."""
        killed = self._walk(g, communicator, cur_step, "timeout")
        return req_send_check, killed
