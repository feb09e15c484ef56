"""Transport interface of the parameter server + the CPU/gloo and single-process tiers.

What the reference does with mpi4py point-to-point + bcast (SURVEY.md §2.2, C1–C10) is expressed as five
operations over flat arenas:

=====================  =================================================  ==============================
operation              reference mechanism                                 here
=====================  =================================================  ==============================
``ps_publish``         isend(step, tag 10) + per-layer bcast (C1, C2)      step word + one weight arena
``worker_fetch``       irecv(tag 10) + per-layer bcast + load_state_dict   wait step word, weights in place
``worker_push``        per-layer isend tag 88+l (C6/C8)                    bucket range of the grad arena
``ps_reduce_update``   Waitany + decompress + sum + /K + opt.step (C5)     K-of-P gather + fused update
``ps_signal_kill`` /   tag 77 Iprobe (C10)                                 abort message / abort word
``worker_poll_kill``
=====================  =================================================  ==============================

``GlooTransport`` is the no-GPU plumbing tier (BASELINE.json config #1).  ``LocalTransport`` is the
single-process tier (PS and one worker share a device; used by N=1 runs and unit tests).  The NVLink
tier lives in ``nvl.py``.
"""
from __future__ import annotations

import time
from typing import List, Optional

import torch
import torch.distributed as dist

from .arena import FlatBuffer, ParamLayout
from ..utils import codec

HDR_OK, HDR_KILLED = 1, 2
TAG_HDR, TAG_KILL, TAG_GRAD, TAG_BUFS = 10, 77, 88, 99  # the reference's tag numbers, kept for the wire-trace's sake


class Transport:
    name = "abstract"

    def __init__(self, rank: int, world_size: int, layout: ParamLayout, colocate: bool = False):
        self.rank, self.world_size, self.layout = rank, world_size, layout
        self.colocate = colocate
        self.is_ps = rank == 0
        self.worker_ranks = list(range(0 if colocate else 1, world_size))
        self.num_workers = len(self.worker_ranks)
        self.stats = {"push_bytes": 0, "bcast_bytes": 0, "stale_dropped": 0, "killed": 0}

    def worker_slot(self, rank: int) -> int:
        return self.worker_ranks.index(rank)

    # PS side
    def ps_publish(self, step: int, weights: FlatBuffer) -> None: raise NotImplementedError
    def ps_reduce_update(self, step: int, k: int, optimizer, master: FlatBuffer, strict_reference=False,
                         kill_stragglers=False, timeout_s: float = 0.0) -> None: raise NotImplementedError
    def ps_signal_kill(self, ranks: List[int], step: int) -> None: pass
    def ps_shutdown(self, step: int) -> None: pass
    # worker side
    def worker_fetch(self, weights: FlatBuffer, timeout_s: float = 0.0, fused_wait: bool = False) -> int: raise NotImplementedError
    def worker_push(self, step: int, grads: FlatBuffer, lo: int = 0, hi: Optional[int] = None, last: bool = True) -> None: raise NotImplementedError
    def worker_push_killed(self, step: int, grads: Optional[FlatBuffer] = None) -> None: pass
    def worker_poll_kill(self, step: int) -> bool: return False
    def mark_pushed(self, b: int) -> None: pass
    def sync_start_step(self, step: Optional[int] = None) -> int: return int(step or 1)      # collective on tiers with device step counters
    def worker_send_buffers(self, step: int, flat: torch.Tensor) -> None: pass
    def ps_recv_buffers(self, step: int, numel: int) -> Optional[torch.Tensor]: return None
    def barrier(self) -> None: pass
    def close(self) -> None: pass


# ---------------------------------------------------------------------------------------------------------
def reduce_k_of_p(arrived: List[torch.Tensor], k: int, strict_reference: bool, owned: bool = False) -> torch.Tensor:
    """Sum the first K arrivals and divide by K.  With ``strict_reference`` the reference's gate quirk is
    reproduced: ``counter <= K`` admits K+1 gradients but the divisor stays K
    (``/root/reference/src/sync_replicas_master_nn.py:179,207``).  ``owned``: the arrivals are private receive buffers, so the
    first one may serve as the accumulator (no copy)."""
    take = min(len(arrived), k + 1 if strict_reference else k)
    acc = arrived[0] if (owned and arrived[0].dtype == torch.float32) else arrived[0].to(torch.float32).clone()
    for g in arrived[1:take]:
        acc.add_(g.to(torch.float32))
    return acc.div_(float(k))


class LocalTransport(Transport):
    """PS and a single worker in one process on one device: pushes are direct arena hand-offs."""
    name = "local"

    def __init__(self, layout: ParamLayout, device="cpu"):
        super().__init__(0, 1, layout, colocate=True)
        self.device = torch.device(device)
        self._step = 0
        self._weights_src: Optional[FlatBuffer] = None
        self._pushed: Optional[torch.Tensor] = None
        self._kill = False

    def ps_publish(self, step, weights):
        self._step, self._weights_src = step, weights
        self.stats["bcast_bytes"] += weights.flat.numel() * weights.flat.element_size()

    def worker_fetch(self, weights, timeout_s=0.0, fused_wait=False):
        if self._weights_src is not None and self._weights_src.flat.data_ptr() != weights.flat.data_ptr():
            weights.flat.copy_(self._weights_src.flat)
        return self._step

    def worker_push(self, step, grads, lo=0, hi=None, last=True):
        self._pushed = grads.flat
        hi = grads.flat.numel() if hi is None else hi
        self.stats["push_bytes"] += (hi - lo) * grads.flat.element_size()

    def ps_reduce_update(self, step, k, optimizer, master, strict_reference=False, kill_stragglers=False, timeout_s=0.0):
        avg = reduce_k_of_p([self._pushed], max(1, min(k, 1)), False)
        optimizer.step(grads=avg)

    def ps_signal_kill(self, ranks, step):
        self._kill = True

    def worker_send_buffers(self, step, flat):
        self._bufs = flat.detach().clone()

    def ps_recv_buffers(self, step, numel):
        return getattr(self, "_bufs", None)

    def worker_poll_kill(self, step):
        k, self._kill = self._kill, False
        return k


class _Rx:
    """Receive state machine for one worker's (header, payload) pair of one step.  gloo's ``Work.is_completed()``
    only turns true inside ``wait()``, so each pair is driven by a small thread and arrival order is observed
    through ``done`` (the ``MPI.Request.Waitany`` role, ``sync_replicas_master_nn.py:160-162``)."""

    def __init__(self, rank: int, step: int):
        import threading
        self.rank, self.step = rank, step
        self.hdr = torch.zeros(2, dtype=torch.int64)
        self.grad_blob = None
        self.done = False
        self._seen = False
        self.error = None
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        try:
            dist.irecv(self.hdr, src=self.rank, tag=self.step * 1000 + TAG_HDR).wait()
            if int(self.hdr[1]) != HDR_KILLED:
                buf = torch.empty(int(self.hdr[0]), dtype=torch.uint8)
                dist.irecv(buf, src=self.rank, tag=self.step * 1000 + TAG_GRAD).wait()
                self.grad_blob = buf
        except Exception as e:  # noqa: BLE001
            self.error = e
        self.done = True

    def poll(self) -> bool:
        """True exactly once, when the pair has just completed."""
        if self.done and not self._seen:
            self._seen = True
            if self.error is not None:
                raise self.error
            return True
        return False


class GlooTransport(Transport):
    """CPU tier over ``torch.distributed`` (gloo).  Weights travel as one collective broadcast of the arena,
    gradients as tagged point-to-point messages (``step*1000 + 88 + bucket`` — the reference's stale-gradient
    tag scheme, ``resnet_split.py:28-42``); the PS consumes them in arrival order and aggregates the first K.
    ``compress=True`` runs the lossless shuffle+LZ codec over every payload (the Blosc/snappy role,
    ``/root/reference/src/compression.py:18-46``)."""
    name = "gloo"

    def __init__(self, layout: ParamLayout, compress: bool = False, colocate: bool = False, group=None):
        assert dist.is_initialized()
        super().__init__(dist.get_rank(), dist.get_world_size(), layout, colocate)
        self.device = torch.device("cpu")
        self.compress = compress
        self.group = group
        self._step_t = torch.zeros(2, dtype=torch.int64)
        self._pending = []     # stale receives still in flight: (work, buffer)
        self._local_push: Optional[torch.Tensor] = None   # colocated worker's gradient
        self._send_reqs = []

    # -- helpers ------------------------------------------------------------------------------------------
    def _bcast_payload(self, t: torch.Tensor, src: int = 0) -> None:
        """Broadcast ``t`` (in place on receivers), optionally through the lossless codec."""
        if not self.compress:
            dist.broadcast(t, src=src, group=self.group)
            self.stats["bcast_bytes"] += t.numel() * t.element_size()
            return
        if self.rank == src:
            blob = codec.compress(t)
            n = torch.tensor([blob.numel()], dtype=torch.int64)
            dist.broadcast(n, src=src, group=self.group)
            dist.broadcast(blob, src=src, group=self.group)
            self.stats["bcast_bytes"] += blob.numel()
        else:
            n = torch.zeros(1, dtype=torch.int64)
            dist.broadcast(n, src=src, group=self.group)
            blob = torch.empty(int(n.item()), dtype=torch.uint8)
            dist.broadcast(blob, src=src, group=self.group)
            codec.decompress_into(blob, t)

    # -- PS -----------------------------------------------------------------------------------------------
    def ps_publish(self, step, weights):
        self._step_t[0], self._step_t[1] = step, 0
        dist.broadcast(self._step_t, src=0, group=self.group)          # C1
        self._bcast_payload(weights.flat, src=0)                        # C2 (one arena, not L messages)
        self._published_weights = weights

    def ps_shutdown(self, step):
        # workers stop on their own at --max-steps (as in the reference); only stragglers' late messages remain
        self._drain(block=True)

    def _drain(self, block: bool):
        """Finish (and drop) receives from stragglers of earlier steps."""
        while True:
            for rx in self._pending:
                if rx.poll() and rx.grad_blob is not None:
                    self.stats["stale_dropped"] += 1
            self._pending = [rx for rx in self._pending if not rx._seen]
            if not block or not self._pending:
                return
            time.sleep(0.0002)

    def ps_reduce_update(self, step, k, optimizer, master, strict_reference=False, kill_stragglers=False, timeout_s=0.0):
        remote = [r for r in self.worker_ranks if r != 0]
        k = max(1, min(k, self.num_workers))
        self._drain(block=True)       # a straggler's previous-step messages are consumed before re-posting
        rxs = [_Rx(r, step) for r in remote]
        arrived: List[torch.Tensor] = []
        if self.colocate and self._local_push is not None:
            arrived.append(self._local_push)
        need = k + 1 if strict_reference else k
        killed_sent = False
        t0 = time.time()
        while True:
            progressed = False
            for rx in rxs:
                if rx.poll():
                    progressed = True
                    if rx.done and rx.grad_blob is not None:
                        if self.compress:
                            g = torch.empty(self.layout.total, dtype=torch.float32)
                            codec.decode_grad(rx.grad_blob, g, True)
                        else:
                            g = rx.grad_blob.view(torch.float32)       # the receive buffer IS the fp32 gradient: no second copy
                            assert g.numel() == self.layout.total
                        arrived.append(g)
                    elif rx.done:
                        self.stats["killed"] += 1
            open_rx = [rx for rx in rxs if not rx._seen]      # 'seen', not 'done': a thread may finish between poll and here
            if not open_rx:
                break
            timed_out = bool(timeout_s) and (time.time() - t0) > timeout_s and len(arrived) >= 1
            if len(arrived) >= need or timed_out:
                if kill_stragglers and not killed_sent:
                    self.ps_signal_kill([rx.rank for rx in open_rx], step)   # C10: stragglers abandon backward
                    killed_sent = True
                if not strict_reference:
                    # first-K semantics: the PS moves on; late messages are received in the background and dropped
                    self._pending += open_rx
                    break
            if not progressed:
                time.sleep(0.0002)
        kk = k if strict_reference else min(k, max(1, len(arrived)))
        avg = reduce_k_of_p(arrived, kk, strict_reference, owned=True)      # receive buffers and the co-located clone are ours to overwrite
        optimizer.step(grads=avg)
        self._local_push = None

    def _store(self):
        from torch.distributed.distributed_c10d import _get_default_store
        return _get_default_store()

    def ps_signal_kill(self, ranks, step):
        """C10: the kill signal is a key in the rendezvous store (the reference polls MPI tag 77 with Iprobe,
        ``resnet_split.py:513-520``; gloo has no probe, and a blocked receive would outlive the step)."""
        st = self._store()
        for r in ranks:
            if r != 0:
                st.set("psb_kill_r%d" % r, str(int(step)))

    # -- worker -------------------------------------------------------------------------------------------
    def worker_fetch(self, weights, timeout_s=0.0, fused_wait=False):
        if self.rank != 0:
            dist.broadcast(self._step_t, src=0, group=self.group)
            if int(self._step_t[1]) == 1:
                return -1                                               # shutdown
            self._bcast_payload(weights.flat, src=0)
        elif getattr(self, "_published_weights", None) is not None and \
                self._published_weights.flat.data_ptr() != weights.flat.data_ptr():
            weights.flat.copy_(self._published_weights.flat)              # co-located worker: local hand-off
        return int(self._step_t[0])

    def worker_push(self, step, grads, lo=0, hi=None, last=True):
        if not last:
            return                                # gloo tier sends the arena once, after the last bucket
        if self.rank == 0:
            self._local_push = grads.flat.detach().to(torch.float32).clone()
            return
        for q in self._send_reqs:
            q.wait()
        payload = codec.encode_grad(grads.flat, self.compress)
        hdr = torch.tensor([payload.numel(), HDR_OK], dtype=torch.int64)
        self._send_reqs = [dist.isend(hdr, dst=0, tag=step * 1000 + TAG_HDR),
                           dist.isend(payload, dst=0, tag=step * 1000 + TAG_GRAD)]
        self._keep = (hdr, payload)
        self.stats["push_bytes"] += payload.numel()

    def worker_push_killed(self, step, grads=None):
        if grads is not None:
            grads.flat.zero_()          # a partial backward must not leak into the next step's accumulation
        if self.rank == 0:
            return
        hdr = torch.tensor([0, HDR_KILLED], dtype=torch.int64)
        self._send_reqs.append(dist.isend(hdr, dst=0, tag=step * 1000 + TAG_HDR))
        self._keep = (hdr,)

    def worker_poll_kill(self, step):
        if self.rank == 0:
            return False
        st, key = self._store(), "psb_kill_r%d" % self.rank
        if st.check([key]):
            return int(st.get(key)) == step
        return False

    def worker_send_buffers(self, step, flat):
        if self.rank == 0:
            self._bufs = flat.detach().float().cpu().clone()
        elif flat.numel():
            dist.send(flat.detach().float().cpu().contiguous(), dst=0, tag=step * 1000 + TAG_BUFS)

    def ps_recv_buffers(self, step, numel):
        if numel == 0:
            return None
        if self.colocate:
            return getattr(self, "_bufs", None)
        buf = torch.empty(numel, dtype=torch.float32)
        dist.recv(buf, src=self.worker_ranks[0], tag=step * 1000 + TAG_BUFS)
        return buf

    def barrier(self):
        dist.barrier(group=self.group)

    def close(self):
        for q in self._send_reqs:
            try:
                q.wait()
            except Exception:
                pass
