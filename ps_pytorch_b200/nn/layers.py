"""``torch.nn``-compatible layers whose CUDA forward/backward run on the sm_100a kernels.

Every class subclasses its ``torch.nn`` namesake, so ``state_dict`` keys and shapes are
those of the reference models (``/root/reference/src/model_ops/*.py``) and checkpoints stay
loadable by either side.  On CPU tensors (the gloo plumbing tier) or with the ``library``
backend the stock ``torch.nn.functional`` op is used.

Fusion points are explicit in the layer API, because on B200 the CNNs here are HBM-bound
outside the convolutions: ``BatchNorm2d(x, residual=..., relu=True)`` is one normalise+add+ReLU
pass, ``MaxPool2d(..., relu=True)`` fuses LeNet's pool→ReLU.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .backend import use_sm100, note_fallback


class Conv2d(nn.Conv2d):
    def forward(self, x, relu: bool = False, grad_holder=None, bn=None):
        """``relu=True`` fuses bias + ReLU into the convolution epilogue (VGG's conv→ReLU pairs).  ``grad_holder`` (``ops.conv.GradHolder``):
        the skip-connection gradient parked there by ``stash_grad`` is added in this convolution's dgrad epilogue.  ``bn``: the
        ``BatchNorm2d`` that consumes this output next; in training its batch statistics are then reduced in the convolution epilogue
        (no separate statistics pass over the activation)."""
        if use_sm100(x):
            from ..ops import conv as _conv
            why = _conv.unsupported_reason(x, self.weight, self.stride, self.padding, self.dilation, self.groups)
            stem = why is not None and _conv.stem_supported(x, self.weight, self.stride, self.padding, self.dilation, self.groups)
            if why is None or stem:
                stats = None
                if bn is not None and not relu and (bn.training or not bn.track_running_stats) and torch.is_grad_enabled():
                    from ..ops import norm as _norm
                    if x.dtype == torch.bfloat16 and bn.num_features == self.out_channels and self.out_channels % 8 == 0 \
                            and (self.out_channels // 8) & (self.out_channels // 8 - 1) == 0 and self.out_channels // 8 <= 256:
                        stats = _conv.stats_buffer(self, self.out_channels, x.device)
                if stem:
                    if grad_holder is not None:
                        grad_holder.consumed = True
                    return _conv.stem_conv2d(x, self.weight, self.bias, relu, stats)
                return _conv.conv2d(x, self.weight, self.bias, self.stride, self.padding, relu, grad_holder, stats)
            note_fallback("conv2d", why)
        if grad_holder is not None:
            grad_holder.consumed = True           # library path: nobody will pick the parked gradient up
        if use_sm100(x):
            from ..ops import libconv as _lc
            if _lc.enabled() and torch.is_grad_enabled() and not isinstance(self.padding, str):
                y = _lc.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
                return F.relu(y) if relu else y
        y = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        return F.relu(y) if relu else y


class Linear(nn.Linear):
    def forward(self, x, relu: bool = False):
        if use_sm100(x):
            from ..ops import gemm as _gemm
            why = _gemm.linear_unsupported_reason(x, self.weight)
            if why is None:
                return _gemm.linear(x, self.weight, self.bias, relu)
            note_fallback("linear", why)
        y = F.linear(x, self.weight, self.bias)
        return F.relu(y) if relu else y


class BatchNorm2d(nn.BatchNorm2d):
    """BatchNorm2d with fused residual-add and ReLU.  Running statistics are kept in fp32
    whatever the parameter dtype (``_apply`` below), so bf16 replicas do not lose the EMA."""

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        for name in ("running_mean", "running_var"):
            buf = getattr(self, name, None)
            if buf is not None and buf.is_floating_point() and buf.dtype != torch.float32:
                setattr(self, name, buf.float())
        return self

    def forward(self, x, residual=None, relu: bool = False):
        if use_sm100(x):
            from ..ops import norm as _norm
            why = _norm.bn_unsupported_reason(x)
            if why is None:
                return _norm.batch_norm(x, self, residual, relu)
            note_fallback("batch_norm", why)
        if self.training and self.track_running_stats and self.num_batches_tracked is not None and not getattr(self, "external_counter", False):
            self.num_batches_tracked.add_(1)
        w, b = self.weight, self.bias
        rm, rv = self.running_mean, self.running_var
        if x.dtype != torch.float32 and x.is_cuda:
            # cuDNN/ATen mixed-precision BN wants fp32 affine params + stats with a 16-bit input
            w = w.float() if w is not None else None
            b = b.float() if b is not None else None
        elif x.dtype != torch.float32:
            # CPU bf16: compute in fp32
            y = F.batch_norm(x.float(), rm, rv, None if w is None else w.float(), None if b is None else b.float(),
                             self.training or not self.track_running_stats, self.momentum, self.eps).to(x.dtype)
            if residual is not None:
                y = y + residual
            return F.relu(y) if relu else y
        y = F.batch_norm(x, rm, rv, w, b, self.training or not self.track_running_stats, self.momentum, self.eps)
        if residual is not None:
            y = y + residual
        return F.relu(y) if relu else y


class ReLU(nn.ReLU):
    def forward(self, x):
        if use_sm100(x):
            from ..ops import elementwise as _ew
            return _ew.relu(x)
        return F.relu(x, inplace=self.inplace)


class MaxPool2d(nn.MaxPool2d):
    def forward(self, x, relu: bool = False):
        if use_sm100(x):
            from ..ops import elementwise as _ew
            why = _ew.pool_unsupported_reason(x, self.kernel_size, self.stride, self.padding)
            if why is None:
                return _ew.max_pool2x2(x, relu)
            if _ew.pool3_supported(x, self.kernel_size, self.stride, self.padding):
                y = _ew.max_pool3x3s2(x)
                return _ew.relu(y) if relu else y
            note_fallback("max_pool2d", why)
        y = F.max_pool2d(x, self.kernel_size, self.stride, self.padding)
        return F.relu(y) if relu else y


class AvgPool2d(nn.AvgPool2d):
    def forward(self, x):
        if use_sm100(x):
            from ..ops import elementwise as _ew
            why = _ew.avgpool_unsupported_reason(x, self.kernel_size)
            if why is None:
                return _ew.global_avg_pool(x)
            note_fallback("avg_pool2d", why)
        return F.avg_pool2d(x, self.kernel_size, self.stride, self.padding)


class Dropout(nn.Dropout):
    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        if use_sm100(x):
            from ..ops import elementwise as _ew
            if _ew.dropout_supported(x, self.p):
                return _ew.dropout(x, self.p, self)
            note_fallback("dropout", "needs a contiguous bf16 tensor with numel % 8 == 0")
        return F.dropout(x, self.p, self.training, self.inplace)


class CrossEntropyLoss(nn.CrossEntropyLoss):
    """Mean softmax cross-entropy (reference: ``nn.CrossEntropyLoss`` at ``distributed_worker.py:98``)."""

    def forward(self, logits, target):
        if use_sm100(logits):
            from ..ops import loss as _loss
            return _loss.softmax_cross_entropy(logits, target)
        return F.cross_entropy(logits.float(), target)
