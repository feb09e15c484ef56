"""BatchNorm2d (+residual, +ReLU) autograd wrapper over ``csrc/bn_kernels.cu``."""
from __future__ import annotations

import os

import torch

from . import count_launch, require

# Measured (profiles/run12.md): the fused last-CTA variant with fp32 atomics is SLOWER on B200 (8.19 vs 7.38 ms/step) — 592 CTAs
# hammering 2*C addresses cost more than the extra launch — so the deterministic two-kernel path is the default.
# Recomputing the ReLU mask from x instead of re-reading the output z saves one streamed tensor per backward pass on paper, but
# measured A/B on B200 (same job, 2 repeats): 7.59 ms/step with it vs 7.44 without — z is still L2-hot from the next layer's
# wgrad and the 16 extra registers cost occupancy.  Off by default.
_MASKX = os.environ.get("PSB200_BN_MASKX", "0") != "0"
_DETERMINISTIC = os.environ.get("PSB200_BN_FUSED_FINALIZE", "0") == "0"
# "wide" finalize kernels (one CTA per 8 channels, all partial loads of a thread in flight at once) are the default since they were
# validated on hardware (6.96 vs 7.23 ms/step, profiles/run24_gated_kernels.md); PSB200_BN_FINALIZE=narrow selects the old ones.
_FINALIZE_NARROW = os.environ.get("PSB200_BN_FINALIZE", "") == "narrow"
_wide_applied = False


def _apply_switches(ext):
    global _wide_applied
    if _FINALIZE_NARROW and not _wide_applied:
        ext.bn_set_finalize_wide(0)
        _wide_applied = True


def _workspace(module, c, device):
    """Self-cleaning per-layer accumulator (2*C sums + ticket) for the fused stats/finalize kernels."""
    if _DETERMINISTIC:
        return None
    ws = getattr(module, "_psb_ws", None)
    if ws is None or ws.device != device or ws.numel() < 2 * c + 1:
        ws = torch.zeros(2 * c + 1, dtype=torch.float32, device=device)
        module._psb_ws = ws
    return ws


def bn_unsupported_reason(x: torch.Tensor):
    if x.dtype != torch.bfloat16:
        return "activations are not bf16"
    if x.dim() != 4:
        return "not a 4-D activation"
    c = x.size(1)
    lc = c // 8
    if c % 8 or lc > 256 or (lc & (lc - 1)):
        return "C/8 is not a power of two <= 256 (C=%d)" % c
    return None


class _BatchNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, training, momentum, eps, relu, ws, conv_partial=None):
        ext = require()
        _apply_switches(ext)
        x = x.contiguous(memory_format=torch.channels_last)
        if residual is not None:
            residual = residual.contiguous(memory_format=torch.channels_last)
        if conv_partial is not None and training:
            # the producing convolution already reduced the batch statistics in its epilogue: finalize + apply only
            z, mean, invstd, ca, cb = ext.bn_forward(x, weight, bias, running_mean, running_var, residual, training, momentum, eps, relu, None,
                                                     conv_partial)
            count_launch(2)
        else:
            z, mean, invstd, ca, cb = ext.bn_forward(x, weight, bias, running_mean, running_var, residual, training, momentum, eps, relu,
                                                     ws if training else None)
            count_launch((2 if ws is not None else 3) if training else 2)
        ctx.relu, ctx.has_res, ctx.training, ctx.ws = relu, residual is not None, training, ws
        # BN+ReLU without a residual: the backward recomputes the ReLU mask from x (a*x+b > 0), so z is neither saved nor re-read
        maskx = relu and residual is None and _MASKX
        ctx.save_for_backward(x, z if (relu and not maskx) else None, weight, mean, invstd, ca if maskx else None, cb if maskx else None)
        return z

    @staticmethod
    def backward(ctx, dz):
        x, z, weight, mean, invstd, ca, cb = ctx.saved_tensors
        if not ctx.training:
            raise RuntimeError("bn backward in eval mode is not supported by the sm100 kernels")
        ext = require()
        dz = dz.contiguous(memory_format=torch.channels_last)
        dx, dres, dgamma, dbeta = ext.bn_backward(dz, z, x, weight, mean, invstd, ctx.relu, ctx.has_res, ctx.ws, ca, cb)
        count_launch(2 if ctx.ws is not None else 3)
        if weight is None:
            dgamma = dbeta = None
        return dx, dgamma, dbeta, (dres if ctx.has_res else None), None, None, None, None, None, None, None, None


def batch_norm(x, module, residual=None, relu=False):
    training = module.training or not module.track_running_stats
    if training and module.track_running_stats and module.num_batches_tracked is not None and not getattr(module, "external_counter", False):
        module.num_batches_tracked.add_(1)
    mom = 0.1 if module.momentum is None else module.momentum
    part = getattr(x, "_psb_stats", None)            # left by the convolution that produced x (ops/conv.py), consumed exactly once
    if part is not None:
        x._psb_stats = None
        part = part.view(-1, 2, x.size(1)) if training else None
    return _BatchNormFn.apply(x, module.weight, module.bias, residual, module.running_mean, module.running_var,
                              training, mom, module.eps, relu, _workspace(module, x.size(1), x.device), part)
