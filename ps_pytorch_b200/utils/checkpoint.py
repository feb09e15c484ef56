"""Checkpoint layout of the reference: ``<train_dir>model_step_<N>`` is ``torch.save(state_dict)``
(``/root/reference/src/sync_replicas_master_nn.py:264-270``, ``distributed_worker.py:301-307``) and is what the
evaluator polls for (``distributed_evaluator.py:80-88``).  New: a single designated writer (fixes the reference's
concurrent-writer race for VGG), atomic rename so the evaluator never sees a torn file, and a
*sidecar* ``model_step_<N>.opt`` with optimizer state + step + RNG so training can resume (the reference cannot)."""
from __future__ import annotations

import os
import re
from typing import Optional

import torch


def model_path(train_dir: str, step: int) -> str:
    # plain concatenation like the reference: train_dir is expected to end in '/'
    return train_dir + "model_step_" + str(step)


def export_state_dict(module: torch.nn.Module, master=None) -> dict:
    """fp32, contiguous, CPU tensors with the reference's keys.  ``master`` (PS fp32 FlatBuffer) overrides the
    parameter values so checkpoints carry master precision rather than the bf16 compute copy."""
    sd = {}
    pnames = [n for n, _ in module.named_parameters()]
    for k, v in module.state_dict().items():
        t = v.detach()
        if master is not None and k in pnames:
            t = master.views[pnames.index(k)].detach()
        if t.is_floating_point():
            t = t.to(torch.float32)
        sd[k] = t.contiguous().clone(memory_format=torch.contiguous_format).cpu()
    return sd


def save_model(module, train_dir: str, step: int, master=None, optimizer=None, extra: Optional[dict] = None) -> str:
    os.makedirs(train_dir, exist_ok=True) if train_dir.endswith("/") else None
    path = model_path(train_dir, step)
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as f:
        torch.save(export_state_dict(module, master), f)
    os.replace(tmp, path)
    if optimizer is not None:
        side = {"step": step, "optimizer": optimizer.flat_state_dict() if hasattr(optimizer, "flat_state_dict") else None,
                "optimizer_sd": optimizer.state_dict() if (not hasattr(optimizer, "flat_state_dict") or optimizer.flat_state_dict() is None) else None,
                "rng": torch.get_rng_state(), "extra": extra or {}}
        tmp = path + ".opt.tmp%d" % os.getpid()
        torch.save(side, tmp)
        os.replace(tmp, path + ".opt")
    return path


def load_model_into(module, path: str, strict: bool = True):
    with open(path, "rb") as f:
        sd = torch.load(f, map_location="cpu")
    own = module.state_dict()
    cast = {k: (v.to(own[k].dtype) if k in own and torch.is_tensor(v) and v.is_floating_point() else v) for k, v in sd.items()}
    module.load_state_dict(cast, strict=strict)
    return module


def latest_step(train_dir: str) -> Optional[int]:
    d = os.path.dirname(train_dir) or "."
    best = None
    if not os.path.isdir(d):
        return None
    for fn in os.listdir(d):
        m = re.fullmatch(r"model_step_(\d+)", fn)
        if m and os.path.isfile(os.path.join(d, fn + ".opt")):
            best = max(best or 0, int(m.group(1)))
    return best


def load_sidecar(train_dir: str, step: int) -> dict:
    return torch.load(model_path(train_dir, step) + ".opt", map_location="cpu", weights_only=False)
