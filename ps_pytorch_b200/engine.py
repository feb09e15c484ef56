"""Programmatic API: one object that is this rank's share of a parameter-server job.

``PSJob`` builds the transport and the roles exactly as ``src/distributed_nn.py`` does from flags, and exposes
``step(x_host, y_host)`` — one synchronous PS iteration on this rank: (rank 0) publish → (workers) wait weights, forward,
backward, push → (server ranks) reduce + optimizer + broadcast.  ``bench.py`` and ``__graft_entry__.smoke`` drive this;
the ``SyncReplicasMaster_NN.start()`` / ``DistributedWorker.train()`` loops of the reference API are the same body in a loop.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import nn as pnn
from .config import num_classes_for
from .models import build_model
from .parallel.arena import ParamLayout
from .parallel.compute import resolve_dtype
from .parallel.ps import SyncReplicasMaster_NN
from .parallel.transport import GlooTransport, LocalTransport
from .parallel.worker import DistributedWorker


class PSJob:
    def __init__(self, network: str = "ResNet18", dataset: str = "Cifar10", batch_size: int = 1024, device="cuda", dtype: str = "auto",
                 lr: float = 0.1, momentum: float = 0.9, optimizer: str = "sgd", weight_decay: float = 0.0, nesterov: bool = False,
                 transport: str = "auto", topology: str = "ps", colocate: bool = True, grad_dtype: str = "bf16", num_aggregate: Optional[int] = None,
                 overlap: bool = False, num_buckets: int = 4, cuda_graph: bool = True, backend: str = "auto", max_steps: int = 10 ** 9,
                 eval_freq: int = 10 ** 9, train_dir: str = "output/models/", mode: str = "normal", compress: bool = False, seed: int = 1,
                 comm_type: str = "Bcast"):
        self.device = torch.device(device)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        if backend != "auto":
            pnn.set_backend(backend)
        torch.manual_seed(seed)
        nc = num_classes_for(dataset)
        self.dtype = resolve_dtype(dtype, self.device)
        proto = build_model(network, nc)
        self.layout = ParamLayout.from_module(proto, channels_last=self.device.type == "cuda")
        n_buf = sum(b.numel() for b in proto.buffers())
        if transport == "auto":
            transport = "nvl" if self.device.type == "cuda" else ("local" if self.world == 1 else "gloo")
        if transport == "local":
            self.comm = LocalTransport(self.layout, self.device)
        elif transport == "gloo":
            self.comm = GlooTransport(self.layout, compress=compress, colocate=colocate)
        else:
            from .parallel.nvl import NvlTransport
            self.comm = NvlTransport(self.layout, self.device, grad_dtype=grad_dtype, colocate=colocate,
                                     num_buckets=num_buckets if overlap else 1, topology=topology, n_buffers=n_buf,
                                     pull_weights=comm_type.lower() == "async")
        self.topology = getattr(self.comm, "topology", "ps")
        P = self.comm.num_workers
        self.k = P if num_aggregate is None else max(1, min(num_aggregate, P))
        if hasattr(self.comm, "follow_server") and (self.k < P or mode != "normal"):
            self.comm.follow_server = True        # a worker can be left behind: its host follows the server's step (see NvlTransport.worker_fetch)
        common = dict(batch_size=batch_size, learning_rate=lr, max_epochs=10 ** 9, momentum=momentum, network=network, eval_freq=eval_freq,
                      train_dir=train_dir, max_steps=max_steps, compress_grad="compress" if compress else "None", device=self.device,
                      mode=mode, dtype=dtype)
        self.master = self.worker = None
        self.is_server = self.rank == 0 or self.topology == "sharded"
        self.is_worker = self.rank in self.comm.worker_ranks
        if self.is_server:
            self.master = SyncReplicasMaster_NN(comm=self.comm, num_aggregate=self.k, optimizer=optimizer, weight_decay=weight_decay,
                                                nesterov=nesterov, worker_dtype=self.dtype, **common)
            self.master.build_model(num_classes=nc, network=proto)
        if self.is_worker:
            self.worker = DistributedWorker(comm=self.comm, overlap=overlap, num_buckets=num_buckets, cuda_graph=cuda_graph,
                                            log_interval=10 ** 9, **common)
            self.worker.build_model(num_classes=nc)
        self.batch_size = batch_size
        self.step_id = 0
        self._first_step = 1              # the step the first step() call computes (resume moves it: sync_start_step)
        self.last_loss = None
        self._comm_events = None          # set by time_comm(True): CUDA-event pairs around the communication kernels of every step

    # ---- exposed communication, measured directly ------------------------------------------------------------------------
    def time_comm(self, on: bool = True) -> None:
        """Bracket the communication kernels of every following step (weights wait, gradient push, PS reduce+update+broadcast) with
        CUDA events on the step's stream.  Those kernels run serialised with the compute on that stream, so the bracketed device time —
        spins included — IS the communication the step is exposed to."""
        self._comm_events = [] if on else None
        if self.worker is not None:        # the weights wait must be a kernel of its own to be bracketed (the one fused into the step then passes at once)
            self.worker.compute.unfused_wait = bool(on)

    def comm_ms_per_step(self) -> Optional[float]:
        """Mean exposed communication per step (ms) on this rank since ``time_comm(True)``; call after a device synchronize."""
        ev = self._comm_events
        if not ev:
            return None
        steps = max(1, len({s for s, _, _ in ev}))
        return sum(a.elapsed_time(b) for _, a, b in ev) / steps

    def _bracket(self):
        if self._comm_events is None or self.device.type != "cuda":
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def _close_bracket(self, e0):
        if e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self._comm_events.append((self.step_id, e0, e1))

    def step(self, x_host: Optional[torch.Tensor] = None, y_host: Optional[torch.Tensor] = None):
        """One PS iteration.  Returns this rank's loss tensor (device, async) or None on a pure-PS rank."""
        self.step_id += 1
        m, w = self.master, self.worker
        if m is not None and (self.rank == 0 or self.topology == "sharded"):
            m.async_bcast_step()
        if w is not None and getattr(self.comm, "follow_server", False) and self.comm.next_worker_step() > self._first_step + self.step_id - 1:
            # K < P: this worker was moved forward to the server's step earlier (it had fallen behind), so it is AHEAD of its call count; computing now
            # would ask for weights the server only publishes in a later call — every rank calls step() the same number of times, so sit this one out
            w = None
        if w is not None:
            e = self._bracket()
            w.sync_fetch_step()
            self._close_bracket(e)
            if w.next_step < 0:
                w = None
        if w is not None:
            w.update_step()
            loss, _ = w.compute.train_step(x_host, y_host, w.cur_step)
            e = self._bracket()
            for b in w.compute.captured_buckets:
                self.comm.mark_pushed(b)
            w._send_grads()
            self._close_bracket(e)
            self.last_loss = loss
        if m is not None:
            e = self._bracket()
            self.comm.ps_reduce_update(m.cur_step, self.k, m.optimizer, m.master, **m._ru_kwargs())
            self._close_bracket(e)
            if m.lowp is not m.master and not getattr(self.comm, "fused_lowp", False):
                m.lowp.flat.copy_(m.master.flat)
            m.cur_step += 1
        return self.last_loss

    def stage(self, x_host: torch.Tensor, y_host: torch.Tensor) -> None:
        """Announce the next step's batch so its host→device copy overlaps the current step (see ``WorkerCompute.stage``)."""
        if self.worker is not None:
            self.worker.compute.stage(x_host, y_host)

    def finish(self) -> None:
        """End of the job on the server side: releases every worker that still waits for a publication (K < P leaves stragglers behind)."""
        if self.master is not None and hasattr(self.comm, "ps_shutdown") and getattr(self.comm, "follow_server", False):
            self.comm.ps_shutdown(self.step_id)

    def close(self):
        self.finish()
        self.comm.close()
