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
        if need_dw or need_db:
            side = _side(dev)
            side.wait_stream(main)                      # dy is complete on the main stream at this point
            with torch.cuda.stream(side):
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
