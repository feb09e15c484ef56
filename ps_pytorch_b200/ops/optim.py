"""Python entry points of the fused flat optimizer kernels (``csrc/optim_kernels.cu``)."""
from __future__ import annotations

from . import count_launch, require


def sgd_flat(w, g, buf, lr, momentum, dampening, weight_decay, nesterov, first, scale=1.0, out_lowp=None):
    count_launch()
    require().sgd_flat(w, g, buf, float(lr), float(momentum), float(dampening), float(weight_decay), bool(nesterov),
                       bool(first), float(scale), out_lowp)


def adam_flat(w, g, m, v, vmax, step_size, beta1, beta2, eps, weight_decay, scale=1.0, out_lowp=None):
    count_launch()
    require().adam_flat(w, g, m, v, vmax, float(step_size), float(beta1), float(beta2), float(eps), float(weight_decay),
                        float(scale), out_lowp)
