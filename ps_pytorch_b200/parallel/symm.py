"""Symmetric peer-mapped arena (python side): allocation through the C++ VMM layer (``ops/csrc/arena.cpp``),
file-descriptor rendezvous over AF_UNIX sockets, multicast (NVLS) binding.

Every rank allocates the same number of bytes; afterwards ``peer_ptrs[r] + off`` addresses rank r's copy of offset
``off`` from any rank's kernels, and ``mc_ptr + off`` (when the box supports multicast) stores to *all* copies at once.
With ``world == 1`` the arena is an ordinary device buffer (single-GPU runs need no mapping)."""
from __future__ import annotations

import os
import socket
import threading
import time
from typing import List

import torch
import torch.distributed as dist

from .. import ops


_job_token = None


def _sock_path(tag: str, rank: int) -> str:
    """Per-JOB rendezvous path: rank 0 draws a random token and broadcasts it over the bootstrap group, so two jobs on one box — even with the
    same (or an unset) MASTER_PORT — never meet in /tmp."""
    global _job_token
    if _job_token is None:
        tok = [os.urandom(6).hex() if (not dist.is_initialized() or dist.get_rank() == 0) else None]
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast_object_list(tok, src=0)
        _job_token = tok[0]
    return "/tmp/psb200_%s_%s_%s_r%d.sock" % (os.environ.get("MASTER_PORT", "0"), _job_token, tag, rank)


class SymmetricArena:
    def __init__(self, nbytes: int, device: torch.device, tag: str = "arena"):
        self.device = torch.device(device)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.mc_ptr = 0
        self._keep = []
        if self.world == 1:
            self.size = (nbytes + 4095) // 4096 * 4096
            self._buf = torch.zeros(self.size, dtype=torch.uint8, device=self.device)
            self.local_ptr = self._buf.data_ptr()
            self.peer_ptrs = [self.local_ptr]
            self.backend = "local"
            return
        which = os.environ.get("PSB200_ARENA", "vmm")
        if which == "torch":
            self._init_torch(nbytes)
        else:
            self._init_vmm(nbytes, tag)
        torch.cuda.synchronize(self.device)
        dist.barrier()

    # ---- own VMM path ---------------------------------------------------------------------------------------
    def _init_vmm(self, nbytes: int, tag: str):
        ext = ops.require()
        torch.cuda.set_device(self.device)
        use_mc = os.environ.get("PSB200_MULTICAST", "1") != "0" and ext.multicast_supported(self.dev_index)
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(use_mc))
        use_mc = all(flags)
        gran = ext.mc_granularity(self.world, max(nbytes, 1 << 21)) if use_mc else 0
        va, handle, size, fd = ext.vmm_alloc(int(nbytes), self.dev_index, int(gran))
        self.size, self.local_ptr, self._handle = size, va, handle
        sizes = [None] * self.world
        dist.all_gather_object(sizes, int(size))
        assert len(set(sizes)) == 1, "symmetric arena sizes differ: %r" % (sizes,)
        mc_handle, mc_fd = 0, -1
        if use_mc and self.rank == 0:
            try:
                mc_handle, mc_fd = ext.mc_create(self.world, size)
            except RuntimeError:
                use_mc = False
        ok = [use_mc]
        dist.broadcast_object_list(ok, src=0)
        use_mc = ok[0]
        # fd rendezvous: every rank serves (its allocation fd [, the multicast fd on rank 0]) to each peer
        path = _sock_path(tag, self.rank)
        if os.path.exists(path):
            os.unlink(path)
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(path)
        srv.listen(self.world)

        def serve():
            for _ in range(self.world - 1):
                c, _a = srv.accept()
                fds = [fd] + ([mc_fd] if (use_mc and self.rank == 0) else [])
                socket.send_fds(c, [b"x"], fds)
                c.recv(1)          # peer has imported; safe to close our end
                c.close()
        th = threading.Thread(target=serve, daemon=True)
        th.start()
        dist.barrier()
        self.peer_ptrs: List[int] = [0] * self.world
        self.peer_ptrs[self.rank] = va
        self._imports = []
        for r in range(self.world):
            if r == self.rank:
                continue
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            for _ in range(200):
                try:
                    c.connect(_sock_path(tag, r))
                    break
                except (FileNotFoundError, ConnectionRefusedError):
                    time.sleep(0.05)
            _msg, fds, _f, _a = socket.recv_fds(c, 1, 2)
            pva, ph = ext.vmm_import(fds[0], size, self.dev_index)
            self.peer_ptrs[r] = pva
            self._imports.append((pva, ph))
            if r == 0 and use_mc and len(fds) > 1:
                mc_handle = ext.mc_import(fds[1])
                os.close(fds[1])
            os.close(fds[0])
            c.send(b"k")
            c.close()
        th.join()
        srv.close()
        os.unlink(path)
        ext.close_fd(fd)
        if mc_fd >= 0:
            ext.close_fd(mc_fd)
        if use_mc:
            ext.mc_add_device(mc_handle, self.dev_index)
            dist.barrier()                    # every device must be added before any bind
            ext.mc_bind(mc_handle, handle, size)
            dist.barrier()
            self.mc_ptr = ext.mc_map(mc_handle, size, self.dev_index, gran)
            self._mc_handle = mc_handle
        self._buf = ext.tensor_from_ptr(va, size, self.dev_index)
        self._buf.zero_()
        self.backend = "vmm+mc" if use_mc else "vmm"

    # ---- torch.distributed._symmetric_memory bring-up shortcut (PSB200_ARENA=torch) ----------------------------
    def _init_torch(self, nbytes: int):
        import torch.distributed._symmetric_memory as symm_mem
        n = (nbytes + 4095) // 4096 * 4096
        t = symm_mem.empty(n, dtype=torch.uint8, device=self.device)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
        self._buf, self._hdl = t, hdl
        self.size = n
        self.local_ptr = t.data_ptr()
        self.peer_ptrs = [int(p) for p in hdl.buffer_ptrs]
        self.mc_ptr = int(hdl.multicast_ptr) if getattr(hdl, "multicast_ptr", 0) else 0
        t.zero_()
        self.backend = "torch-symm" + ("+mc" if self.mc_ptr else "")

    def close(self):
        """Collective: unmap the multicast view (unbind + release the multicast object), the peer mappings and the local allocation."""
        if getattr(self, "backend", "").startswith("vmm") and not getattr(self, "_closed", False):
            torch.cuda.synchronize(self.device)
            if dist.is_initialized():
                dist.barrier()
            ext = ops.require()
            if self.mc_ptr:
                try:
                    ext.mc_unmap(self.mc_ptr, getattr(self, "_mc_handle", 0), self.size, self.dev_index)
                    self.mc_ptr = 0
                except RuntimeError:
                    pass                      # an old driver refusing the unbind: keep the mapping (leaks VA until process exit)
                if dist.is_initialized():
                    dist.barrier()            # nobody unmaps memory that a peer's multicast object still references
            for va, h in self._imports:
                ext.vmm_unmap(va, h, self.size)
            if not self.mc_ptr:
                self._buf = None
                ext.vmm_unmap(self.local_ptr, self._handle, self.size)
            self._closed = True

    # ---- views ------------------------------------------------------------------------------------------------
    def view(self, offset: int, nbytes: int, dtype: torch.dtype) -> torch.Tensor:
        assert offset % 16 == 0 and offset + nbytes <= self.size
        return self._buf[offset: offset + nbytes].view(dtype)

    def peer_view(self, rank: int, offset: int, nbytes: int, dtype: torch.dtype) -> torch.Tensor:
        """A tensor over rank ``rank``'s copy (for bulk copies through cudaMemcpy; kernels use raw pointers)."""
        if rank == self.rank or self.world == 1:
            return self.view(offset, nbytes, dtype)
        t = ops.require().tensor_from_ptr(self.peer_ptrs[rank] + offset, nbytes, self.dev_index)
        return t.view(dtype)
