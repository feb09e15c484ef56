"""Kernel backend selection for ``ps_pytorch_b200.nn`` layers.

``sm100``   — hand-written sm_100a kernels from ``ps_pytorch_b200.ops`` (product path on B200).
``library`` — cuDNN/cuBLAS/ATen through stock ``torch.nn.functional`` (CPU tier and the baseline arm).
"""
from __future__ import annotations

import contextlib
import logging
import threading

import torch

_state = threading.local()
_default = "auto"
_warned = set()
log = logging.getLogger("ps_pytorch_b200.nn")


def set_backend(name: str) -> None:
    global _default
    assert name in ("auto", "sm100", "library"), name
    _default = name


def get_backend() -> str:
    return getattr(_state, "override", None) or _default


def use_sm100(x: torch.Tensor) -> bool:
    """True when ``x`` should run through the hand-written kernels."""
    if not x.is_cuda:
        return False
    b = get_backend()
    if b == "library":
        return False
    from .. import ops
    if b == "sm100":
        ops.require()  # fail loudly if the extension is missing on a GPU box
        return True
    return ops.available()


@contextlib.contextmanager
def backend(name: str):
    prev = getattr(_state, "override", None)
    _state.override = name
    try:
        yield
    finally:
        _state.override = prev


def note_fallback(op: str, why: str) -> None:
    """Library fallbacks inside the sm100 backend are never silent: logged once per (op, reason)."""
    key = (op, why)
    if key not in _warned:
        _warned.add(key)
        log.warning("sm100 backend: %s uses the library path (%s)", op, why)
