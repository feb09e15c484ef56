"""Library (cuDNN) convolution with the weight gradient on a side stream.

cuDNN's ``convolution_backward`` runs dgrad and wgrad back to back on one stream.  In a CNN backward pass the next thing on
the critical path after dgrad is a memory-bound BatchNorm backward, while wgrad is tensor-bound and nobody needs its result
until the gradients are pushed — so here dgrad stays on the main stream and wgrad is issued on a side stream (fork after
``dy`` is ready, join before the gradients are gathered; inside a CUDA-graph capture this becomes a fork/join in the graph).
``PSB200_WGRAD_STREAM=0`` keeps everything on one stream."""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

_ENABLED = os.environ.get("PSB200_WGRAD_STREAM", "1") != "0"
# dedicated stem kernel: correct but 199 us vs cuDNN 189 us at batch 1024 (scripts/bench_stem.py), and hidden on the side stream: opt-in
_STEM = os.environ.get("PSB200_STEM_WGRAD", "0") != "0"
_stem_scratch = {}


def _stem_ok(x, weight, stride, padding, dilation, groups, dy):
    return (_STEM and weight.size(1) <= 3 and tuple(weight.shape[2:]) == (3, 3) and tuple(stride) == (1, 1) and tuple(padding) == (1, 1)
            and tuple(dilation) == (1, 1) and groups == 1 and weight.size(0) % 64 == 0 and x.size(2) % 8 == 0 and x.size(3) <= 32
            and x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last))
_streams = {}
_forked = {}


def enabled() -> bool:
    return _ENABLED


def _side(device):
    s = _streams.get(device)
    if s is None:
        s = _streams[device] = torch.cuda.Stream(device=device)
    return s


def join(device) -> None:
    """Make the current stream wait for outstanding side-stream wgrads (call before reading parameter gradients)."""
    if _forked.get(device):
        torch.cuda.current_stream(device).wait_stream(_streams[device])
        _forked[device] = False


class _LibConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups):
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, padding, dilation, groups, bias is not None)
        return F.conv2d(x, weight, bias, stride, padding, dilation, groups)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        stride, padding, dilation, groups, has_bias = ctx.cfg
        need_dx, need_dw, need_db = ctx.needs_input_grad[0], ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]
        dx = dw = db = None
        dev = dy.device
        main = torch.cuda.current_stream(dev)
        if (need_dw or need_db) and (weight.grad is not None):
            # a bound .grad is accumulated on the main stream as soon as backward returns: no side stream in that mode
            join(dev)                                   # (a shared weight's first gradient may still be in flight on the side stream)
            _, dw, db = torch.ops.aten.convolution_backward(dy, x, weight, [weight.size(0)] if need_db else None, stride, padding, dilation,
                                                            False, [0, 0], groups, [False, need_dw, need_db])
        elif need_dw or need_db:
            side = _side(dev)
            side.wait_stream(main)                      # dy is complete on the main stream at this point
            with torch.cuda.stream(side):
                if need_dw and not need_db and _stem_ok(x, weight, stride, padding, dilation, groups, dy):
                    # network stem (RGB -> 64): dedicated register-blocked kernel instead of cuDNN's generic implicit GEMM
                    from . import count_launch, require
                    key = (dev, tuple(weight.shape))
                    sc = _stem_scratch.get(key)
                    if sc is None:
                        sc = _stem_scratch[key] = torch.zeros(weight.numel(), dtype=torch.float32, device=dev)
                    dw = require().stem_wgrad(dy.contiguous(memory_format=torch.channels_last), x, sc)
                    count_launch(2)
                else:
                    _, dw, db = torch.ops.aten.convolution_backward(dy, x, weight, [weight.size(0)] if need_db else None, stride, padding, dilation,
                                                                    False, [0, 0], groups, [False, need_dw, need_db])
            _forked[dev] = True
            for t in (dy, x, weight):
                t.record_stream(side)
        if need_dx:
            dx, _, _ = torch.ops.aten.convolution_backward(dy, x, weight, None, stride, padding, dilation, False, [0, 0], groups, [True, False, False])
        return dx, dw, db, None, None, None, None


def conv2d(x, weight, bias, stride, padding, dilation, groups):
    return _LibConvFn.apply(x, weight, bias, tuple(stride), tuple(padding), tuple(dilation), groups)
