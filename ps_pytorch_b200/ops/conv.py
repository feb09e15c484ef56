"""Convolutions on the tcgen05 implicit-GEMM kernels (``csrc/gemm_tcgen05.cu``)."""
from __future__ import annotations

import torch

from . import count_launch, require


def unsupported_reason(x, weight, stride, padding, dilation, groups):
    ext = require()
    if not hasattr(ext, "conv_fprop"):
        return "tcgen05 implicit-GEMM conv not built"
    return ext.conv_unsupported_reason(list(x.shape), list(weight.shape), list(stride), list(padding), list(dilation), groups,
                                       x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16)


def conv2d(x, weight, bias, stride, padding):
    from .conv_impl import conv2d as _c
    return _c(x, weight, bias, stride, padding)
