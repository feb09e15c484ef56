"""Fused softmax-cross-entropy (loss + dL/dlogits in one pass) over ``csrc/elementwise.cu``."""
from __future__ import annotations

import torch

from . import count_launch, require


class _SoftmaxCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        loss, dl = require().softmax_ce(logits.contiguous(), target.contiguous())
        count_launch(2)
        ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g.to(dl.dtype), None


def softmax_cross_entropy(logits, target):
    return _SoftmaxCEFn.apply(logits, target)
