"""Convolutions on the hand-written tcgen05 implicit-GEMM kernels (``csrc/gemm_tcgen05.cu``, ``csrc/conv2_tcgen05.cu``).

Routing (all three passes are ours; cuDNN is only reachable through ``PSB200_TC_CONV=0`` / ``set_tc_conv(False)``):

===========================  =====================================================================================================
forward                      3x3/stride-1 maps with W >= 8, H >= 16: CTA-pair *halo* kernel (input patch fetched once per channel block,
                             taps are shifted smem windows); everything else (strided, 1x1, small maps): CTA-pair *tap-table* kernel
data gradient, stride 1      the same two kernels on the transposed (+180-degree rotated for the halo form) filter
data gradient, stride 2      tap-table kernel, one parity class of output pixels per tile group — no zero insertion
weight gradient              split-K tcgen05 kernel with MN-major operands (fp32 vector reductions into a scratch)
===========================  =====================================================================================================

``PSB200_TC_CONV``: "1" (default) every supported convolution on our kernels; "0" the library path (A/B runs)."""
from __future__ import annotations

import os

import torch

from . import count_launch, require
from ..nn.backend import note_fallback

_MODE = os.environ.get("PSB200_TC_CONV", "1").lower()
_ENABLED = _MODE != "0"
_applied = False


def set_tc_conv(on) -> None:
    global _ENABLED, _MODE
    _MODE = on if isinstance(on, str) else ("1" if on else "0")
    _ENABLED = _MODE != "0"


def tc_conv_enabled() -> bool:
    return _ENABLED


def _apply_switches(ext):
    global _applied
    if _applied:
        return
    if os.environ.get("PSB200_TC_HALO_PAIR", "") != "":
        ext.tc_set_halo_pair(int(os.environ["PSB200_TC_HALO_PAIR"]))
    if os.environ.get("PSB200_TC_HALO_PITCH", "") != "":           # 16 = 2 KB patch rows, 0 = tight (8 + S - 1 pixels)
        ext.tc_set_halo_pitch(int(os.environ["PSB200_TC_HALO_PITCH"]))
    if os.environ.get("PSB200_WGRAD_SMS", "") != "":               # cap the grid of the side-stream weight-gradient kernels (experiments)
        ext.tc_set_wgrad_sms(int(os.environ["PSB200_WGRAD_SMS"]))
    if os.environ.get("PSB200_TC_CONV_ALGO", "") != "":            # -1 auto (halo where eligible), 0 per-tap, 1 halo
        ext.tc_set_conv_algo(int(os.environ["PSB200_TC_CONV_ALGO"]))
    _applied = True


def unsupported_reason(x, weight, stride, padding, dilation, groups):
    if not _ENABLED:
        return "tcgen05 conv disabled (PSB200_TC_CONV=0)"
    ext = require()
    if not hasattr(ext, "conv2"):
        return "tcgen05 implicit-GEMM conv not built"
    _apply_switches(ext)
    if isinstance(padding, str):
        return "string padding"
    r = ext.conv_unsupported_reason(list(x.shape), list(weight.shape), list(stride), list(padding), list(dilation), groups,
                                    x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16)
    if r:
        return r
    if weight.size(0) % 64:
        return "Cout is not a multiple of 64"
    if stride[0] not in (1, 2):
        return "stride > 2"
    if weight.size(2) * weight.size(3) > 9:
        return "filter larger than 9 taps"
    if stride[0] == 2 and (x.size(2) % 2 or x.size(3) % 2):
        return "strided convolution of an odd-sized map"
    return None


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


def _halo_ok(h_out, w_out, r, s, stride):
    return stride == 1 and r * s > 1 and s <= 8 and w_out >= 8 and h_out >= 16


def fprop(ext, x, w, bias, stride, pad, relu, residual=None, stats=None):
    r, s = w.size(2), w.size(3)
    ho, wo = (x.size(2) + 2 * pad - r) // stride + 1, (x.size(3) + 2 * pad - s) // stride + 1
    count_launch()
    if _halo_ok(ho, wo, r, s, stride):
        return ext.conv_fprop(x, w, bias, stride, pad, relu, residual, stats)
    return ext.conv2(x, w, bias, stride, pad, relu, residual, False, 0, 0, stats)


_STATS_IN_EPILOGUE = os.environ.get("PSB200_CONV_BN_STATS", "1") != "0"


def stats_buffer(module, cout, device):
    """Persistent fp32 [4 * #SMs][2][Cout] partial-statistics rows written by the conv epilogue (one private row per epilogue warp) and
    read + re-zeroed by the BatchNorm finalize kernel; None when the fusion is switched off (PSB200_CONV_BN_STATS=0)."""
    if not _STATS_IN_EPILOGUE:
        return None
    buf = getattr(module, "_psb_stats_buf", None)
    if buf is None or buf.device != device:
        buf = torch.zeros(4 * require().num_sms() * 2 * cout, dtype=torch.float32, device=device)
        module._psb_stats_buf = buf
    return buf


# ---- dgrad operands ---------------------------------------------------------------------------------------------------------------
# The data gradient needs the filter as [Cin][R][S][Cout] (rotated by 180 degrees for the halo form).  Transposing per layer costs a
# 4-9 us launch in front of every dgrad (measured: d_ours vs d_pair_kernel_only, profiles/run25_conv_per_layer.md).  A training loop that
# owns the step (WorkerCompute) instead opts in to ONE batched transpose of all registered weights at the start of each step:
# enable_batched_dgrad_operands(True) + refresh_dgrad_operands() before every backward.
_wt_reg = {}            # id(weight) -> [weight, flip, buf, data_ptr]
_wt_plan = None         # (table, n, blocks, number of registered entries)
_wt_batched = False
_wt_fresh = False


def enable_batched_dgrad_operands(on: bool) -> None:
    global _wt_batched, _wt_fresh, _wt_plan
    _wt_batched, _wt_fresh = bool(on), False
    _wt_reg.clear()                  # a new owner (one model per process): forget the previous model's filters
    _wt_plan = None


def refresh_ready() -> bool:
    """True when the next ``refresh_dgrad_operands`` call will launch the batched kernel (so it can carry the weights-ready wait)."""
    return bool(_wt_batched and _wt_reg)


def refresh_dgrad_operands(wait=None) -> bool:
    """One launch: re-transpose every registered conv weight (call once per step, after the weights of the step are in place).
    ``wait = (flags_ptr, nflags, step_ptr, err_ptr, timeout_s)``: the kernel first acquires the transport's "weights of this step published"
    words — it is the first reader of the weights in the step, so the separate wait kernel is not needed.  Returns True iff the wait rode along."""
    global _wt_plan, _wt_fresh
    if not _wt_batched or not _wt_reg:
        _wt_fresh = False
        return False
    ext = require()
    ents = list(_wt_reg.values())
    if _wt_plan is None or _wt_plan[3] != len(ents) or any(e[0].data_ptr() != e[3] for e in ents):
        for e in ents:
            e[3] = e[0].data_ptr()
        table, meta = ext.weight_transpose_plan([e[0].detach() for e in ents], [e[2] for e in ents], [int(e[1]) for e in ents])
        _wt_plan = (table, int(meta[0]), int(meta[1]), len(ents))
    if wait is not None:
        ext.weight_transpose_run(_wt_plan[0], _wt_plan[1], _wt_plan[2], *wait)
    else:
        ext.weight_transpose_run(_wt_plan[0], _wt_plan[1], _wt_plan[2])
    count_launch()
    _wt_fresh = True
    return wait is not None


def _dgrad_operand(ext, weight, flip: bool):
    ent = _wt_reg.get(id(weight)) if _wt_batched else None
    if ent is not None and ent[1] == flip and _wt_fresh and _wt_plan is not None and _wt_plan[3] == len(_wt_reg) and ent[3] == weight.data_ptr():
        return ent[2]
    w = _krsc(weight)
    count_launch()
    out = ext.weight_transpose(w.detach(), flip)
    if _wt_batched and w is weight and weight.is_leaf:          # arena-backed parameter with a stable address: batch it from the next step on
        _wt_reg[id(weight)] = [weight, flip, out.clone() if ent is None else ent[2], weight.data_ptr()]
    return out


def dgrad(ext, dy, weight, in_h, in_w, stride, pad, residual=None):
    """dX of conv(x, weight, stride, pad) from dY; ``residual`` (shape of dX) is added in the kernel epilogue."""
    r, s = weight.size(2), weight.size(3)
    count_launch()
    if stride == 1 and _halo_ok(in_h, in_w, r, s, 1) and r == s and in_h == dy.size(2) + r - 1 - 2 * pad:
        # dX = conv(dY, rot180(W)^T, pad = R-1-pad): the halo kernel on the transformed filter
        return ext.conv_fprop(dy, _dgrad_operand(ext, weight, True), None, 1, r - 1 - pad, False, residual)
    return ext.conv2(dy, _dgrad_operand(ext, weight, False), None, stride, pad, False, residual, True, in_h, in_w)


class GradHolder:
    """Rendezvous for the skip-connection gradient of a residual block (see ``stash_grad``)."""
    __slots__ = ("grad", "consumed")

    def __init__(self):
        self.grad, self.consumed = None, False


class _StashGrad(torch.autograd.Function):
    """Identity in forward.  In backward the incoming gradient is parked in the holder instead of being returned, and the convolution that
    shares this input adds it in its dgrad epilogue (``residual=``): autograd's separate accumulate kernel at every skip connection
    (8 per ResNet-18 step) disappears.  If that convolution has already run (or never will), the gradient is returned normally."""

    @staticmethod
    def forward(ctx, x, holder):
        ctx.holder = holder
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        h = ctx.holder
        if h.consumed:
            return g, None
        h.grad = g if h.grad is None else h.grad + g
        return None, None


def stash_grad(x, holder):
    return _StashGrad.apply(x, holder)


skip_fused_count = 0             # how many skip gradients were added inside a dgrad epilogue (tests assert the fusion really happens)
_SKIP_FUSION = os.environ.get("PSB200_SKIP_GRAD_FUSION", "1") != "0"
_WGRAD_SIDE = os.environ.get("PSB200_TC_WGRAD_STREAM", "1") != "0"


# ---- gradient sinks: in gather mode (p.grad is None) the weight-gradient kernels write straight into the transport's gradient arena --------
_grad_sinks = {}                 # weight.data_ptr() -> channels_last bf16 view of the gradient arena
_sinks_active = False            # only inside WorkerCompute's step body: a stray backward elsewhere must get its gradient the normal way
sink_writes = 0                  # weight gradients that landed in a sink (tests assert the copy-free path really runs)


def set_grad_sinks(params, views) -> list:
    """Register arena views as the destination of the weight gradients of `params` (4-D bf16 CUDA weights whose view is channels_last like the
    kernels' output).  Returns the indices that got a sink; their ``.grad`` stays None after backward — read the arena instead."""
    _grad_sinks.clear()
    idx = []
    for i, (p, v) in enumerate(zip(params, views)):
        if (p.dim() == 4 and p.is_cuda and p.dtype == torch.bfloat16 and v.dtype == torch.bfloat16 and v.shape == p.shape
                and v.is_contiguous(memory_format=torch.channels_last)):
            _grad_sinks[p.data_ptr()] = v
            idx.append(i)
    return idx


_sinks_written = set()           # data_ptr of every weight whose sink was written since the last activation


def activate_grad_sinks(on: bool) -> None:
    global _sinks_active
    _sinks_active = bool(on)
    if on:
        _sinks_written.clear()


def sinks_written() -> set:
    """Weights (by data_ptr) whose gradient already sits in its sink for the current step."""
    return _sinks_written


def _sink_for(weight):
    if not _sinks_active or weight.grad is not None:
        return None
    ptr = weight.data_ptr()
    v = _grad_sinks.get(ptr)
    if v is None or ptr in _sinks_written:
        # a weight used twice in one forward (shared layer): the second gradient goes the ordinary way and the gather ADDS it to the sink
        return None
    _sinks_written.add(ptr)
    return v


def set_skip_fusion(on: bool) -> None:
    global _SKIP_FUSION
    _SKIP_FUSION = bool(on)


def new_grad_holder(x):
    """A holder when fusing the skip gradient is possible for this input (our conv path, training), else None."""
    if not (_ENABLED and _SKIP_FUSION and torch.is_grad_enabled() and x.is_cuda and x.requires_grad and x.dtype == torch.bfloat16):
        return None
    return GradHolder()


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, relu, holder=None, stats=None):
        ext = require()
        x = x.contiguous(memory_format=torch.channels_last)
        y = fprop(ext, x, _krsc(weight), bias, stride, pad, relu, None, stats)
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.stride, ctx.pad, ctx.has_bias, ctx.relu, ctx.holder = stride, pad, bias is not None, relu, holder
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
        h = ctx.holder
        skip = None
        if h is not None:
            if need_dx and h.grad is not None and h.grad.shape == x.shape and h.grad.dtype == dy.dtype:
                skip = h.grad.contiguous(memory_format=torch.channels_last)
                h.grad = None
                global skip_fused_count
                skip_fused_count += 1
            else:
                h.consumed = True                     # from now on the stash hands its gradient to autograd itself
        if need_dx:
            dx = dgrad(ext, dy, weight, x.size(2), x.size(3), s, p, residual=skip)
            if h is not None and h.grad is not None:  # parked but not fusable: add the slow way
                dx = dx + h.grad.to(dx.dtype)
                h.grad = None
        if need_dw and weight.size(1) % 64 == 0:
            # The weight gradient is tensor-bound and nobody needs it until the gradients are pushed, while the next thing on the critical
            # path after dgrad is a memory-bound BatchNorm backward: issue wgrad on a side stream (fork after dy is ready; WorkerCompute joins
            # before it gathers the gradients: libconv.join).  Inside a CUDA-graph capture this becomes a fork/join in the graph.
            from . import libconv as _lc
            # Only when autograd will merely HAND OVER the gradient (weight.grad is None: gather mode).  With a bound .grad (overlap / DDP modes)
            # autograd accumulates `grad += dw` on the main stream right after this function returns — before a side-stream kernel has run.
            sink = _sink_for(weight)
            if _WGRAD_SIDE and _lc.enabled() and weight.grad is None:
                dev = dy.device
                main, side = torch.cuda.current_stream(dev), _lc._side(dev)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    dw = ext.conv_wgrad(dy, x, _wgrad_scratch(weight), r, weight.size(3), s, p, sink)
                _lc._forked[dev] = True
                for t in (dy, x, dw):
                    t.record_stream(side)
            else:
                if weight.grad is not None:
                    _lc.join(dy.device)   # shared weight: the gradient autograd is about to add to may still be in flight on the side stream
                dw = ext.conv_wgrad(dy, x, _wgrad_scratch(weight), r, weight.size(3), s, p, sink)
            if sink is not None:          # the gradient is already where the push kernel reads it: nothing to hand to autograd
                global sink_writes
                sink_writes += 1
                dw = None
            count_launch(2)
            need_dw = False
        if need_dw:
            note_fallback("conv2d.backward", "weight gradient with Cin % 64 != 0 uses cuDNN")
            _gi, dw, _ = torch.ops.aten.convolution_backward(dy, x, weight, None, (s, s), (p, p), (1, 1), False, (0, 0), 1, (False, True, False))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum((0, 2, 3)).to(dy.dtype)
        return dx, dw, db, None, None, None, None, None


# ---- network stem (Cin <= 3): im2col rows built by threads, tensor-core forward + weight gradient (csrc/stem_tcgen05.cu) ---------------
def stem_supported(x, weight, stride, padding, dilation, groups) -> bool:
    if not _ENABLED or x.requires_grad or groups != 1 or isinstance(padding, str):
        return False
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16 or tuple(dilation) != (1, 1):
        return False
    ext = require()
    return hasattr(ext, "stem_conv") and bool(ext.stem_eligible(x.size(2), x.size(3), x.size(1), weight.size(0), weight.size(2), weight.size(3),
                                                                int(stride[0]) if stride[0] == stride[1] else 0, int(padding[0]) if padding[0] == padding[1] else -1))


class _StemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu, stats):
        ext = require()
        x = x.contiguous(memory_format=torch.channels_last)
        y = ext.stem_conv(x, _krsc(weight), bias, relu, stats)
        count_launch()
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.relu, ctx.has_bias = relu, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        ext = require()
        dy = dy.contiguous(memory_format=torch.channels_last)
        if ctx.relu:
            dy = ext.relu_backward(dy, y)
            count_launch()
        dw = db = None
        if ctx.needs_input_grad[1]:
            sink = _sink_for(weight)
            dw = ext.stem_wgrad_tc(dy, x, _wgrad_scratch(weight), sink)
            if sink is not None:
                global sink_writes
                sink_writes += 1
                dw = None
            count_launch(2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum((0, 2, 3)).to(dy.dtype)
        return None, dw, db, None, None


def stem_conv2d(x, weight, bias, relu=False, stats=None):
    y = _StemFn.apply(x, weight, bias, bool(relu), stats)
    if stats is not None:
        y._psb_stats = stats
    return y


def conv2d(x, weight, bias, stride, padding, relu=False, grad_holder=None, stats=None):
    y = _ConvFn.apply(x, weight, bias, int(stride[0]), int(padding[0]), bool(relu), grad_holder, stats)
    if stats is not None:
        y._psb_stats = stats          # picked up (once) by the BatchNorm that consumes y: ops/norm.py:batch_norm
    return y
