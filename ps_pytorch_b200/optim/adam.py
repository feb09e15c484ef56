"""PS-side Adam/AMSGrad with ``step(grads)`` (reference ``/root/reference/src/optim/adam.py:12-94``):
L2 weight decay added to the gradient, ``m = β1 m + (1-β1) g``, ``v = β2 v + (1-β2) g²``, optional AMSGrad max,
``denom = sqrt(v) + eps``, ``step_size = lr·sqrt(1-β2^t)/(1-β1^t)``, ``p -= step_size · m/denom``.
The reference imports it on the PS but never constructs it; here ``--optimizer adam`` wires it in."""
from __future__ import annotations

import math

import torch
from torch.optim.optimizer import Optimizer

from ._common import as_tensor_list


class Adam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid betas: {}".format(betas))
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        self._flat = None

    def attach_flat(self, master) -> None:
        z = lambda: torch.zeros_like(master.flat)
        self._flat = {"master": master, "m": z(), "v": z(), "steps": 0,
                      "vmax": z() if self.param_groups[0]['amsgrad'] else None}

    def flat_state(self):
        return self._flat

    @torch.no_grad()
    def step(self, grads=None, closure=None, scale: float = 1.0, out_lowp: torch.Tensor = None):
        loss = closure() if closure is not None else None
        group = self.param_groups[0]
        if grads is None:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in group['params']]
        glist, gflat = as_tensor_list(grads, group['params'])
        if gflat is not None:
            assert self._flat is not None, "flat gradients need attach_flat()"
            self._step_flat(gflat, group, scale, out_lowp)
            return loss
        for g in self.param_groups:
            b1, b2 = g['betas']
            for i, p in enumerate(g['params']):
                grad = glist[i].to(torch.float32)
                if scale != 1.0:
                    grad = grad * scale
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p.data, dtype=torch.float32)
                    st['exp_avg_sq'] = torch.zeros_like(p.data, dtype=torch.float32)
                    if g['amsgrad']:
                        st['max_exp_avg_sq'] = torch.zeros_like(p.data, dtype=torch.float32)
                st['step'] += 1
                if g['weight_decay'] != 0:
                    grad = grad.add(p.data.float(), alpha=g['weight_decay'])
                st['exp_avg'].mul_(b1).add_(grad, alpha=1 - b1)
                st['exp_avg_sq'].mul_(b2).addcmul_(grad, grad, value=1 - b2)
                if g['amsgrad']:
                    torch.maximum(st['max_exp_avg_sq'], st['exp_avg_sq'], out=st['max_exp_avg_sq'])
                    denom = st['max_exp_avg_sq'].sqrt().add_(g['eps'])
                else:
                    denom = st['exp_avg_sq'].sqrt().add_(g['eps'])
                step_size = g['lr'] * math.sqrt(1 - b2 ** st['step']) / (1 - b1 ** st['step'])
                p.data.add_((st['exp_avg'] / denom).to(p.dtype), alpha=-step_size)
        return loss

    def _step_flat(self, gflat, group, scale, out_lowp):
        fs = self._flat
        fs["steps"] += 1
        t = fs["steps"]
        b1, b2 = group['betas']
        step_size = group['lr'] * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        w = fs["master"].flat
        if w.is_cuda:
            from ..ops import optim as _k
            _k.adam_flat(w, gflat, fs["m"], fs["v"], fs["vmax"], step_size, b1, b2, group['eps'],
                         group['weight_decay'], scale, out_lowp)
        else:
            g = gflat.to(torch.float32) * scale
            if group['weight_decay'] != 0:
                g = g.add(w, alpha=group['weight_decay'])
            fs["m"].mul_(b1).add_(g, alpha=1 - b1)
            fs["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
            if fs["vmax"] is not None:
                torch.maximum(fs["vmax"], fs["v"], out=fs["vmax"])
                denom = fs["vmax"].sqrt().add_(group['eps'])
            else:
                denom = fs["v"].sqrt().add_(group['eps'])
            w.addcdiv_(fs["m"], denom, value=-step_size)
            if out_lowp is not None:
                out_lowp.copy_(w)

    def flat_state_dict(self):
        if self._flat is None:
            return None
        fs = self._flat
        return {"kind": "adam", "m": fs["m"].cpu(), "v": fs["v"].cpu(), "steps": fs["steps"],
                "vmax": None if fs["vmax"] is None else fs["vmax"].cpu()}

    def load_flat_state_dict(self, sd):
        fs = self._flat
        fs["m"].copy_(sd["m"]); fs["v"].copy_(sd["v"]); fs["steps"] = int(sd["steps"])
        if fs["vmax"] is not None and sd.get("vmax") is not None:
            fs["vmax"].copy_(sd["vmax"])
