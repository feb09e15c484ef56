"""PS-side SGD whose ``step`` takes the aggregated gradients as an argument.

Semantics are those of the reference (``/root/reference/src/optim/sgd.py:59-92``): ``d_p = grad (+ wd·p)``;
momentum buffer is ``d_p`` on the first step and ``momentum·buf + (1-dampening)·d_p`` afterwards;
Nesterov uses ``d_p + momentum·buf``; ``p -= lr·d_p``.  On a CUDA flat arena the whole model is updated by
one fused kernel (``ops.optim.sgd_flat``) that can also emit the bf16 compute copy.
"""
from __future__ import annotations

import torch
from torch.optim.optimizer import Optimizer

from ._common import as_tensor_list


class SGD(Optimizer):
    def __init__(self, params, lr, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        if lr is None or lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if momentum < 0.0:
            raise ValueError("Invalid momentum value: {}".format(momentum))
        if weight_decay < 0.0:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening,
                                      weight_decay=weight_decay, nesterov=nesterov))
        self._flat = None  # (master FlatBuffer, momentum flat tensor, step counter)

    # ---- flat-arena fast path --------------------------------------------------------------------
    def attach_flat(self, master) -> None:
        """Bind to a FlatBuffer holding every parameter (fp32 master on the PS)."""
        self._flat = {"master": master, "buf": torch.zeros_like(master.flat), "steps": 0}

    def flat_state(self):
        return self._flat

    @torch.no_grad()
    def step(self, grads=None, closure=None, scale: float = 1.0, out_lowp: torch.Tensor = None):
        loss = closure() if closure is not None else None
        group = self.param_groups[0]
        if grads is None:  # torch.optim-style call
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in group['params']]
        glist, gflat = as_tensor_list(grads, group['params'])
        if gflat is not None:
            assert self._flat is not None, "flat gradients need attach_flat()"
            self._step_flat(gflat, group, scale, out_lowp)
            return loss
        for g in self.param_groups:
            wd, mom, damp, nest, lr = g['weight_decay'], g['momentum'], g['dampening'], g['nesterov'], g['lr']
            for i, p in enumerate(g['params']):
                d_p = glist[i].to(torch.float32)
                if scale != 1.0:
                    d_p = d_p * scale
                pf = p.data
                if wd != 0:
                    d_p = d_p.add(pf.float(), alpha=wd)
                if mom != 0:
                    st = self.state[p]
                    if 'momentum_buffer' not in st:
                        buf = st['momentum_buffer'] = d_p.clone()
                    else:
                        buf = st['momentum_buffer']
                        buf.mul_(mom).add_(d_p, alpha=1 - damp)
                    d_p = d_p.add(buf, alpha=mom) if nest else buf
                pf.add_(d_p.to(pf.dtype), alpha=-lr)
        return loss

    def _step_flat(self, gflat, group, scale, out_lowp):
        fs = self._flat
        w = fs["master"].flat
        first = fs["steps"] == 0
        if w.is_cuda:
            from ..ops import optim as _k
            _k.sgd_flat(w, gflat, fs["buf"], group['lr'], group['momentum'], group['dampening'],
                        group['weight_decay'], group['nesterov'], first, scale, out_lowp)
        else:
            d_p = gflat.to(torch.float32) * scale
            if group['weight_decay'] != 0:
                d_p = d_p.add(w, alpha=group['weight_decay'])
            if group['momentum'] != 0:
                buf = fs["buf"]
                if first:
                    buf.copy_(d_p)
                else:
                    buf.mul_(group['momentum']).add_(d_p, alpha=1 - group['dampening'])
                d_p = d_p.add(buf, alpha=group['momentum']) if group['nesterov'] else buf
            w.add_(d_p, alpha=-group['lr'])
            if out_lowp is not None:
                out_lowp.copy_(w)
        fs["steps"] += 1

    # ---- checkpoint sidecar ----------------------------------------------------------------------
    def flat_state_dict(self):
        if self._flat is None:
            return None
        return {"kind": "sgd", "buf": self._flat["buf"].detach().cpu(), "steps": self._flat["steps"]}

    def load_flat_state_dict(self, sd):
        self._flat["buf"].copy_(sd["buf"])
        self._flat["steps"] = int(sd["steps"])
