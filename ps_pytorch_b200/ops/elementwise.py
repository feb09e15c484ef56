"""ReLU / MaxPool2d(2,2)(+ReLU) / global average pool autograd wrappers over ``csrc/elementwise.cu``."""
from __future__ import annotations

import torch

from . import count_launch, require


def _as_pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class _ReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = require().relu_forward(x)
        count_launch()
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        if dy.stride() != y.stride():
            dy = dy.contiguous(memory_format=torch.channels_last) if y.dim() == 4 and y.is_contiguous(memory_format=torch.channels_last) else dy.contiguous()
        count_launch()
        return require().relu_backward(dy, y)


def relu(x):
    dense = x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last))
    if x.dtype != torch.bfloat16 or x.numel() % 8 or not dense:
        return torch.relu(x)
    return _ReluFn.apply(x)


def pool_unsupported_reason(x, kernel_size, stride, padding):
    if _as_pair(kernel_size) != (2, 2) or _as_pair(stride if stride is not None else kernel_size) != (2, 2) or _as_pair(padding) != (0, 0):
        return "only 2x2/2 pooling"
    if x.dtype != torch.bfloat16 or x.dim() != 4 or x.size(1) % 8 or x.size(2) % 2 or x.size(3) % 2:
        return "needs bf16 NHWC, C%8==0, even H/W"
    return None


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, relu_):
        x = x.contiguous(memory_format=torch.channels_last)
        y = require().maxpool2x2_forward(x, relu_)
        count_launch()
        ctx.relu = relu_
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        count_launch()
        return require().maxpool2x2_backward(dy.contiguous(memory_format=torch.channels_last), x, ctx.relu), None


def max_pool2x2(x, relu_=False):
    return _MaxPoolFn.apply(x, relu_)


def avgpool_unsupported_reason(x, kernel_size):
    k = _as_pair(kernel_size)
    if x.dim() != 4 or (x.size(2), x.size(3)) != k:
        return "only global pooling (kernel == spatial size)"
    if x.dtype != torch.bfloat16 or x.size(1) % 8:
        return "needs bf16 NHWC, C%8==0"
    return None


class _GapFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous(memory_format=torch.channels_last)
        ctx.hw = (x.size(2), x.size(3))
        count_launch()
        return require().gap_forward(x)

    @staticmethod
    def backward(ctx, dy):
        count_launch()
        return require().gap_backward(dy, ctx.hw[0], ctx.hw[1])


def global_avg_pool(x):
    return _GapFn.apply(x)
