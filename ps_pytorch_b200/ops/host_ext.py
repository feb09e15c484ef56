"""Loader for the CPU-only C++ extension (codec + batch prefetcher)."""
from __future__ import annotations

_mod = None
_tried = False


def get():
    global _mod, _tried
    if _mod is None and not _tried:
        _tried = True
        try:
            from . import _load_so
            _mod = _load_so("ps_b200_host")
        except Exception:
            _mod = None
    return _mod
