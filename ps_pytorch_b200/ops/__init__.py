"""Native extensions.

``_C``       — CUDA kernels for sm_100a (``csrc/*.cu``): comm (push / reduce-update / flags), fused optimizers,
               BN / pool / loss, tcgen05 GEMM + implicit-GEMM convolution, plus the VMM symmetric arena (``arena.cpp``).
``host_ext`` — CPU-only C++ (``csrc/host_ext.cpp``): lossless shuffle+LZ codec, threaded batch prefetcher.

Both are built in-tree by ``build.py`` (``__graft_entry__.build()``) into ``ops/_build/`` so the ``.so`` files travel
with the repository snapshot to the GPU box.  ``require()`` raises if the CUDA extension is missing: on a GPU box
the sm100 backend never silently falls back to PyTorch.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(_HERE, "_build")
_lock = threading.Lock()
_mod = None
_err = None
launch_count = 0      # incremented by every python wrapper that launches one of our kernels


def count_launch(n: int = 1) -> None:
    global launch_count
    launch_count += n


def _load_so(name: str):
    path = os.path.join(BUILD_DIR, name + ".so")
    if not os.path.isfile(path):
        raise ImportError("%s not built (run `python __graft_entry__.py build`)" % path)
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


def get():
    """The CUDA extension module, or raises ImportError."""
    global _mod, _err
    if _mod is not None:
        return _mod
    with _lock:
        if _mod is None:
            try:
                _mod = _load_so("ps_b200_C")
                if os.environ.get("PSB200_PDL", "0") == "1" and hasattr(_mod, "set_pdl"):
                    _mod.set_pdl(1)            # programmatic dependent launch of the step's kernels (csrc/pdl.cuh): measured slower, opt-in
            except Exception as e:  # noqa: BLE001
                _err = e
                raise
    return _mod


def available() -> bool:
    import torch
    if not torch.cuda.is_available():
        return False
    try:
        get()
        return True
    except Exception:
        return False


def require():
    try:
        return get()
    except Exception as e:
        raise RuntimeError("ps_pytorch_b200: the sm_100a extension is required on a GPU box but failed to load: %r. "
                           "Build it with `python __graft_entry__.py build`." % (e,)) from e
