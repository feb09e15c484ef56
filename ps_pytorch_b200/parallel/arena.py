"""Flat parameter / gradient arenas.

The reference moves one tagged message per parameter tensor (62 for ResNet-18, 161 for ResNet-50;
``/root/reference/src/sync_replicas_master_nn.py:218-237``, ``distributed_worker.py:254-272``).  Here every
parameter, gradient, momentum buffer and wire slot is a *view into one contiguous arena* described by a
single offset table, so a push / reduce-update / broadcast is one kernel over one address range and
per-layer granularity survives only as flag granularity (buckets).

Convolution weights are stored physically as [Cout][kh][kw][Cin] (channels-last) so the same bytes are
the K-major B operand of the implicit-GEMM kernels; their logical shape stays [Cout,Cin,kh,kw], which
keeps ``state_dict`` compatibility.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch
import torch.nn as nn

ALIGN = 128  # elements; every tensor starts on a 256 B (bf16) / 512 B (fp32) boundary: TMA/vector friendly


@dataclass
class ParamLayout:
    names: List[str]
    shapes: List[Tuple[int, ...]]
    numels: List[int]
    offsets: List[int]
    channels_last: List[bool]
    total: int            # padded element count of the arena

    @staticmethod
    def from_module(module: nn.Module, channels_last: bool = True, chunk: int = 1 << 16) -> "ParamLayout":
        names, shapes, numels, offsets, cl = [], [], [], [], []
        off = 0
        for n, p in module.named_parameters():
            names.append(n)
            shapes.append(tuple(p.shape))
            numels.append(p.numel())
            offsets.append(off)
            cl.append(bool(channels_last and p.dim() == 4))
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        total = max(chunk, (off + chunk - 1) // chunk * chunk)
        return ParamLayout(names, shapes, numels, offsets, cl, total)

    @property
    def num_params(self) -> int:
        return sum(self.numels)

    def index_range(self, indices: Sequence[int]) -> Tuple[int, int]:
        """Smallest [lo, hi) element range covering the given parameter indices (hi is ALIGN-padded)."""
        lo = min(self.offsets[i] for i in indices)
        hi = max(self.offsets[i] + (self.numels[i] + ALIGN - 1) // ALIGN * ALIGN for i in indices)
        return lo, hi

    def buckets(self, n: int) -> List[Tuple[int, int, List[int]]]:
        """Split parameters (forward order) into ``n`` contiguous buckets of roughly equal size.
        Returns [(lo, hi, param_indices)], used as the flag granularity for overlapped pushes."""
        n = max(1, min(n, len(self.numels)))
        target = self.num_params / n
        out, cur, acc = [], [], 0
        for i, ne in enumerate(self.numels):
            cur.append(i)
            acc += ne
            if acc >= target * (len(out) + 1) and len(out) < n - 1:
                out.append(cur)
                cur = []
        if cur:
            out.append(cur)
        res = []
        for k, idx in enumerate(out):
            lo = self.offsets[idx[0]]
            hi = self.offsets[out[k + 1][0]] if k + 1 < len(out) else self.total
            res.append((lo, hi, idx))
        return res


class FlatBuffer:
    """One contiguous tensor + per-parameter views in the layout's logical shapes."""

    def __init__(self, layout: ParamLayout, dtype: torch.dtype, device, storage: torch.Tensor = None):
        self.layout = layout
        if storage is None:
            storage = torch.zeros(layout.total, dtype=dtype, device=device)
        assert storage.numel() >= layout.total and storage.dtype == dtype
        self.flat = storage[: layout.total]
        self.views: List[torch.Tensor] = []
        for shp, ne, off, cl in zip(layout.shapes, layout.numels, layout.offsets, layout.channels_last):
            v = self.flat[off: off + ne]
            if cl:
                co, ci, kh, kw = shp
                v = v.view(co, kh, kw, ci).permute(0, 3, 1, 2)
            else:
                v = v.view(shp)
            self.views.append(v)

    @property
    def dtype(self):
        return self.flat.dtype

    @property
    def device(self):
        return self.flat.device

    def copy_from_module(self, module: nn.Module) -> None:
        with torch.no_grad():
            for v, p in zip(self.views, module.parameters()):
                v.copy_(p.detach())

    def copy_to_module(self, module: nn.Module) -> None:
        with torch.no_grad():
            for v, p in zip(self.views, module.parameters()):
                p.copy_(v)

    def bind_params(self, module: nn.Module, keep_arena_values: bool = False) -> None:
        """Make the module's parameters views of this arena.  By default the module's values move into the arena;
        with ``keep_arena_values`` the arena content wins (a replica the PS has written or will write)."""
        with torch.no_grad():
            for v, p in zip(self.views, module.parameters()):
                if not keep_arena_values:
                    v.copy_(p.detach())
                p.data = v

    def bind_grads(self, module: nn.Module) -> None:
        """Make ``p.grad`` views of this arena so backward accumulates straight into the push source."""
        for v, p in zip(self.views, module.parameters()):
            p.grad = v

    def zero_(self):
        self.flat.zero_()
        return self

    def to_list(self, clone: bool = False):
        return [v.clone() if clone else v for v in self.views]

    def load_list(self, tensors) -> None:
        with torch.no_grad():
            for v, t in zip(self.views, tensors):
                v.copy_(torch.as_tensor(t).reshape(v.shape))


def module_device(module: nn.Module):
    return next(module.parameters()).device
