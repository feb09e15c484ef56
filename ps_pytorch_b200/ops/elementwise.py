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


class _MaxPool3Fn(torch.autograd.Function):
    """MaxPool2d(3, stride 2, padding 1) — the ImageNet ResNet stem."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous(memory_format=torch.channels_last)
        y = require().maxpool3x3s2_forward(x)
        count_launch()
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        count_launch()
        return require().maxpool3x3s2_backward(dy.contiguous(memory_format=torch.channels_last), x, y)


def pool3_supported(x, kernel_size, stride, padding) -> bool:
    return (_as_pair(kernel_size) == (3, 3) and _as_pair(stride if stride is not None else kernel_size) == (2, 2) and _as_pair(padding) == (1, 1)
            and x.dtype == torch.bfloat16 and x.dim() == 4 and x.size(1) % 8 == 0)


def max_pool3x3s2(x):
    return _MaxPool3Fn.apply(x)


# ---- dropout: counter-based mask recomputed in backward; the seed word lives on the device so CUDA graphs replay with fresh masks ----
_drop_state = {}
_drop_salt = [0]


def dropout_state(device):
    st = _drop_state.get(device)
    if st is None:
        st = _drop_state[device] = torch.full((1,), int(torch.initial_seed()) & 0x7FFFFFFF, dtype=torch.int64, device=device)
    return st


def bump_dropout_seed(device) -> None:
    """Advance the device-resident dropout seed (once per training step; WorkerCompute does it inside the captured step)."""
    if device in _drop_state:
        require().bump_word(_drop_state[device])
        count_launch()


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, salt):
        st = dropout_state(x.device)
        y, used = require().dropout_forward(x, p, st, salt)
        count_launch()
        ctx.p = p
        ctx.save_for_backward(used)
        return y

    @staticmethod
    def backward(ctx, dy):
        (used,) = ctx.saved_tensors
        count_launch()
        return require().dropout_backward(dy.contiguous(), ctx.p, used), None, None


def dropout_supported(x, p) -> bool:
    return x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() % 8 == 0 and 0.0 < p < 1.0


def dropout(x, p, module=None):
    """Each Dropout module gets a fixed salt (so two layers never share a mask); the per-step variation comes from the device seed word."""
    salt = getattr(module, "_psb_salt", None) if module is not None else None
    if salt is None:
        _drop_salt[0] += 1
        salt = _drop_salt[0] * 0x5851F42D
        if module is not None:
            module._psb_salt = salt
    return _DropoutFn.apply(x, float(p), int(salt))


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
