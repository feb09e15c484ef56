"""Top-k precision, one definition instead of the reference's four verbatim copies
(``/root/reference/src/nn_ops.py:14-27`` et al.)."""
from __future__ import annotations

import torch


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)):
    """Computes the precision@k for the specified values of k (percent, 1-element tensors)."""
    maxk = min(max(topk), output.size(1))
    batch_size = target.size(0)
    _, pred = output.float().topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand(maxk, batch_size))
    res = []
    for k in topk:
        kk = min(k, maxk)
        res.append(correct[:kk].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / batch_size))
    return res
