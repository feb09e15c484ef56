"""Linear layers on the tcgen05 GEMM (``csrc/gemm_tcgen05.cu``)."""
from __future__ import annotations

import torch

from . import count_launch, require


def linear_unsupported_reason(x, weight):
    ext = require()
    if not hasattr(ext, "gemm_tn"):
        return "tcgen05 GEMM not built"
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        return "needs bf16 operands"
    if x.dim() != 2:
        return "needs a 2-D input"
    k = x.size(1)
    if k % 8:
        return "K must be a multiple of 8 (TMA 16-byte row pitch)"
    return None


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        ext = require()
        x = x.contiguous()
        y = ext.gemm_tn(x, weight, bias, relu)           # y[M,N] = x[M,K] @ weight[N,K]^T (+bias)(+relu)
        count_launch()
        ctx.relu = relu
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        ext = require()
        dy = dy.contiguous()
        if ctx.relu:
            dy = dy * (y > 0).to(dy.dtype)
        dx = dw = db = None
        n, m = dy.size(1), dy.size(0)
        if ctx.needs_input_grad[0]:
            # dx[M,K] = dy[M,N] @ W[N,K]  -> TN form needs W^T [K,N] K-major: small matrices, transpose once
            if n % 8 == 0:
                dx = ext.gemm_tn(dy, weight.t().contiguous(), None, False)
                count_launch()
            else:      # reduction dim (e.g. 10 classes) is not a TMA-legal row pitch: zero-pad it to a multiple of 8
                pad = (-n) % 8
                dx = ext.gemm_tn(torch.nn.functional.pad(dy, (0, pad)), torch.nn.functional.pad(weight.t(), (0, pad)).contiguous(), None, False)
                count_launch()
        if ctx.needs_input_grad[1]:
            if m % 8 == 0:
                dw = ext.gemm_tn(dy.t().contiguous(), x.t().contiguous(), None, False)   # [N,M] @ [K,M]^T
            else:
                pad = (-m) % 8
                dw = ext.gemm_tn(torch.nn.functional.pad(dy.t(), (0, pad)).contiguous(), torch.nn.functional.pad(x.t(), (0, pad)).contiguous(), None, False)
            count_launch()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum(0).to(dy.dtype)
        return dx, dw, db, None


def linear(x, weight, bias=None, relu=False):
    return _LinearFn.apply(x, weight, bias, relu)
