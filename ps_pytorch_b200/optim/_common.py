from __future__ import annotations

import numpy as np
import torch


def as_tensor_list(grads, params):
    """Accept what the reference's ``step(grads)`` accepts — a list of numpy arrays aligned with the
    parameters (``/root/reference/src/optim/sgd.py:75-76``) — plus torch tensors and a flat arena."""
    if isinstance(grads, torch.Tensor):
        return None, grads
    out = []
    for g, p in zip(grads, params):
        if isinstance(g, np.ndarray):
            g = torch.from_numpy(np.ascontiguousarray(g))
        out.append(g.to(device=p.device).reshape(p.shape))
    return out, None
