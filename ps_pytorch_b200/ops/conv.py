"""Convolutions on the tcgen05 implicit-GEMM kernels (``csrc/gemm_tcgen05.cu``).

Forward (fprop: per-tap / halo / weight-stationary variants), the data gradient of stride-1 convolutions (the same kernels on
flipped+transposed weights) and the weight gradient (MN-major operands, split-K with fp32 vector reductions) run on the
hand-written kernels; only the data gradient of *strided* convolutions still goes through cuDNN
(``torch.ops.aten.convolution_backward``), logged once per process by ``note_fallback``."""
from __future__ import annotations

import os

import torch

from . import count_launch, require
from ..nn.backend import note_fallback

# Measured on B200 (profiles/run4_conv_microbench.md) the per-tap tcgen05 kernel reaches 0.38-0.76 PFLOP/s where cuDNN's
# sm100 kernels reach 0.75-1.34, so the *default* convolution path is the library one and the hand-written kernel is opt-in
# (PSB200_TC_CONV=1 / set_tc_conv(True)) until it wins; every other layer of the sm100 backend is ours regardless.
_MODE = os.environ.get("PSB200_TC_CONV", "0").lower()       # "0" library, "1" every supported conv, "auto" only where ours wins
_ENABLED = _MODE != "0"
_HALO_PAIR = os.environ.get("PSB200_TC_HALO_PAIR", "0") != "0"   # EXPERIMENTAL cta_group::2 halo kernel, not yet run on hardware
_pair_applied = False


def set_tc_conv(on) -> None:
    global _ENABLED, _MODE
    _MODE = on if isinstance(on, str) else ("1" if on else "0")
    _ENABLED = _MODE != "0"


def tc_conv_enabled() -> bool:
    return _ENABLED


def unsupported_reason(x, weight, stride, padding, dilation, groups):
    if not _ENABLED:
        return "tcgen05 conv is opt-in (PSB200_TC_CONV=1); cuDNN is faster today"
    ext = require()
    if not hasattr(ext, "conv_fprop"):
        return "tcgen05 implicit-GEMM conv not built"
    global _pair_applied
    if not _pair_applied:
        if _HALO_PAIR:
            ext.tc_set_halo_pair(1)
        if os.environ.get("PSB200_TC_CONV_ALGO", "") != "":            # -1 auto (per-tap), 0 per-tap, 1 halo where eligible
            ext.tc_set_conv_algo(int(os.environ["PSB200_TC_CONV_ALGO"]))
        _pair_applied = True
    if isinstance(padding, str):
        return "string padding"
    if _MODE == "auto":
        # measured policy (profiles/run20_gemm.md): 1x1 stride-1 convolutions with >= 256 output channels are plain GEMMs on which the
        # 256-column tcgen05 kernel matches or beats the library
        if not (weight.size(2) == 1 and weight.size(3) == 1 and stride[0] == 1 and weight.size(0) % 256 == 0):
            return "auto policy keeps this shape on cuDNN"
    r = ext.conv_unsupported_reason(list(x.shape), list(weight.shape), list(stride), list(padding), list(dilation), groups,
                                    x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16)
    return r or None


_scratch = {}


def _wgrad_scratch(weight):
    """Persistent, self-cleaning fp32 accumulation buffer for the split-K weight gradient of this weight shape."""
    key = (weight.device, tuple(weight.shape))
    buf = _scratch.get(key)
    if buf is None:
        buf = _scratch[key] = torch.zeros(weight.numel(), dtype=torch.float32, device=weight.device)
    return buf


def _krsc(w):
    """[Cout,Cin,R,S] logical -> physically [Cout][R][S][Cin] (no copy for channels-last parameters)."""
    return w if w.is_contiguous(memory_format=torch.channels_last) else w.contiguous(memory_format=torch.channels_last)


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, relu):
        ext = require()
        x = x.contiguous(memory_format=torch.channels_last)
        y = ext.conv_fprop(x, _krsc(weight), bias, stride, pad, relu)
        count_launch()
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.stride, ctx.pad, ctx.has_bias, ctx.relu = stride, pad, bias is not None, relu
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        ext = require()
        dy = dy.contiguous(memory_format=torch.channels_last)
        if ctx.relu:
            dy = ext.relu_backward(dy, y)
            count_launch()
        s, p = ctx.stride, ctx.pad
        r = weight.size(2)
        dx = dw = db = None
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if need_dx and s == 1 and weight.size(0) % 64 == 0 and weight.size(1) % 8 == 0:
            # dX = conv(dY, flip(W)^T, pad = R-1-pad): the same implicit-GEMM kernel on transformed weights
            wt = weight.flip(2, 3).permute(1, 0, 2, 3).contiguous(memory_format=torch.channels_last)
            dx = ext.conv_fprop(dy, wt, None, 1, r - 1 - p, False)
            count_launch()
            need_dx = False
        if need_dw and weight.size(0) % 64 == 0 and weight.size(1) % 64 == 0:
            dw = ext.conv_wgrad(dy, x, _wgrad_scratch(weight), r, weight.size(3), s, p)
            count_launch(2)
            need_dw = False
        if need_dx or need_dw:
            note_fallback("conv2d.backward", "strided data gradient uses cuDNN")
            gi, gw, _ = torch.ops.aten.convolution_backward(dy, x, weight, None, (s, s), (p, p), (1, 1), False, (0, 0), 1,
                                                            (need_dx, need_dw, False))
            if need_dx:
                dx = gi
            if need_dw:
                dw = gw
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum((0, 2, 3)).to(dy.dtype)
        return dx, dw, db, None, None, None


def conv2d(x, weight, bias, stride, padding, relu=False):
    return _ConvFn.apply(x, weight, bias, int(stride[0]), int(padding[0]), bool(relu))
