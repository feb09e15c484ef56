"""All-reduce data parallelism — the capability of the reference's vendored (and unused) ``DistributedDataParallel``
(``/root/reference/src/data_parallel_dist/data_parallel_dist.py:29-267``: broadcast parameters from rank 0, bucketed
``all_reduce`` of gradients after backward, user-owned optimizer).

Here the module's parameters and gradients are views of flat arenas, and the gradient all-reduce is ONE kernel per rank over
peer-mapped memory (``reduce_update_kernel`` with the optimizer switched off, ``kind = 2``): every rank owns 1/W of the arena,
pulls that shard from all peers with P2P loads, averages in fp32 and stores the result back into every peer's gradient arena
(reduce-scatter + all-gather fused, in place).  On CPU tensors it is a gloo ``all_reduce`` of the flat arena.

    ddp = DistributedDataParallel(model, device)        # broadcasts rank 0's parameters
    loss = criterion(ddp(x), y); loss.backward()
    ddp.sync_gradients()                                # averaged gradients in p.grad on every rank
    optimizer.step()
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from .arena import FlatBuffer, ParamLayout


class DistributedDataParallel(nn.Module):
    def __init__(self, module: nn.Module, device=None, broadcast_buffers: bool = True):
        super().__init__()
        self.module = module
        dev = torch.device(device) if device is not None else next(module.parameters()).device
        self.device = dev
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.broadcast_buffers = broadcast_buffers
        cl = dev.type == "cuda"
        self.module.to(dev)
        if cl:
            self.module.to(memory_format=torch.channels_last)
        self.layout = ParamLayout.from_module(self.module, channels_last=cl)
        dtype = next(self.module.parameters()).dtype
        self._nvl = None
        if dev.type == "cuda" and self.world > 1 and dtype == torch.bfloat16:
            from .nvl import NvlTransport
            self._nvl = NvlTransport(self.layout, dev, grad_dtype="bf16", topology="sharded")
            wstore = self._nvl.worker_weight_storage(torch.bfloat16)
            gstore = self._nvl.worker_grad_storage(torch.bfloat16)
        else:
            wstore = gstore = None
        self.weights = FlatBuffer(self.layout, dtype, dev, storage=wstore)
        self.weights.bind_params(self.module)
        self.grads = FlatBuffer(self.layout, dtype, dev, storage=gstore)
        self.grads.bind_grads(self.module)
        self._sync_params()

    def _sync_params(self):
        """Parameters (and buffers) of rank 0 become everyone's (reference ``:46,128-141``)."""
        if self.world == 1:
            return
        if self._nvl is not None:
            self._nvl.ps_publish(1, self.weights)          # rank 0 multicasts its arena into every replica
            self._nvl.worker_fetch(self.weights)
            torch.cuda.synchronize(self.device)
        else:
            t = self.weights.flat if self.device.type == "cpu" else self.weights.flat.cpu()
            dist.broadcast(t, src=0)
            if self.device.type != "cpu":
                self.weights.flat.copy_(t)
        if self.broadcast_buffers:
            for b in self.module.buffers():
                t = b.detach().cpu()
                dist.broadcast(t, src=0)
                b.data.copy_(t)

    def forward(self, *a, **kw):
        return self.module(*a, **kw)

    def zero_grad(self, set_to_none: bool = False):
        self.grads.flat.zero_()

    @torch.no_grad()
    def sync_gradients(self):
        """Average gradients over ranks, in place in the gradient arena."""
        if self.world == 1:
            return
        if self._nvl is not None:
            self._nvl.all_reduce_grads(self.grads)
        else:
            t = self.grads.flat if self.device.type == "cpu" else self.grads.flat.float().cpu()
            dist.all_reduce(t)
            t.div_(self.world)
            if self.device.type != "cpu":
                self.grads.flat.copy_(t)
