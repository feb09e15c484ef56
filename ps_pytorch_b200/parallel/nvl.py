"""NVLink tier of the parameter server: gradient push, PS reduce+update and weight broadcast are fused kernels over
peer-mapped memory (``ops/csrc/comm_kernels.cu``); ``torch.distributed`` is only the bootstrap.

Two topologies share the same kernels:

``ps``       the reference topology (rank 0 = parameter server).  Workers *push*: cast/quantise + peer-store into their
             slot on the PS + release flag.  The PS kernel polls flags (first K of P), reduces the K slots from its own
             HBM, applies the optimizer to the fp32 master and stores the bf16 weights into every worker replica
             (one NVLS multicast store when available) + releases their "weights ready" flags.
``sharded``  every rank is a worker *and* the server of 1/W of the arena (the role the vendored DDP all-reduce plays in the
             reference, ``/root/reference/src/data_parallel_dist/data_parallel_dist.py:252-261``, without the single-PS
             incast): after backward a rank only raises its "gradients ready" flag on every peer; each owner *pulls* its
             shard from all peers' gradient arenas with P2P loads, updates, and stores the new weights to all replicas.
             Reduce-scatter + optimizer + all-gather in one kernel per rank.

All step counters are device-resident, so the whole worker step and the whole PS step are graph-capturable and the host
never blocks on a peer; every device-side wait has a timeout that sets ``err`` instead of hanging the GPU.
"""
from __future__ import annotations

import os
import time
import torch
import torch.distributed as dist

from .arena import FlatBuffer, ParamLayout
from .symm import SymmetricArena
from .transport import Transport
from .. import ops

WIRE = {"bf16": 0, "fp8": 1, "fp32": 2}
MAXB = 16          # kMaxBuckets in comm_kernels.cu
MAXP = 8
DONE_STEP = 1 << 40  # "training is over": published as a weights-ready value by ps_shutdown; every waiter passes, every follower stops


def _align(x, a=4096):
    return (x + a - 1) // a * a


class NvlTransport(Transport):
    name = "nvl"
    fused_lowp = True            # the reduce-update kernel writes the low-precision replicas itself

    def __init__(self, layout: ParamLayout, device, grad_dtype: str = "bf16", colocate: bool = False, num_buckets: int = 1,
                 topology: str = "ps", n_buffers: int = 0, timeout_s: float = None, weight_dtype: torch.dtype = torch.bfloat16,
                 pull_weights: bool = False):
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        if world == 1:
            colocate = True
        if topology == "sharded":
            colocate = True
        super().__init__(rank, world, layout, colocate)
        assert world <= MAXP, "one NVSwitch box: at most 8 ranks"
        assert weight_dtype == torch.bfloat16, "NVLink tier broadcasts bf16 replicas"
        self.device = torch.device(device)
        self.topology = topology
        self.wire = WIRE[grad_dtype]
        self.grad_dtype = grad_dtype
        self.timeout_s = float(os.environ.get("PSB200_TIMEOUT_S", timeout_s or 120.0))
        self.ext = ops.require()
        T = layout.total
        self.T = T
        P = self.num_workers
        if topology == "ps":
            self.buckets = [(lo, hi) for lo, hi, _ in layout.buckets(min(num_buckets, MAXB))]
            self.bucket_params = [idx for _, _, idx in layout.buckets(min(num_buckets, MAXB))]
        else:
            # shard boundaries on 128-element multiples; rank r owns [cuts[r], cuts[r+1])
            per = (T // world) // 128 * 128
            cuts = [r * per for r in range(world)] + [T]
            self.shards = [(cuts[r], cuts[r + 1]) for r in range(world)]
            self.buckets = [(0, T)]
            self.bucket_params = [list(range(len(layout.numels)))]
        self.NB = len(self.buckets)
        # ---- arena layout (identical on every rank) --------------------------------------------------------
        self.off_w = 0
        self.off_flags = _align(T * 2)
        self.off_grads = self.off_flags + 4096
        gbytes = T * 4                                   # room for an fp32 gradient arena
        self.off_slots = _align(self.off_grads + gbytes)
        self.slot_q_bytes = {0: T * 2, 1: T, 2: T * 4}[self.wire]
        self.slot_s_bytes = _align(T // 32, 256) if self.wire == 1 else 0
        self.slot_stride = _align(self.slot_q_bytes + self.slot_s_bytes, 256)
        nslots = P if topology == "ps" else 0
        self.off_bufs = _align(self.off_slots + nslots * self.slot_stride)
        self.n_buffers = int(n_buffers)
        total = self.off_bufs + _align(max(self.n_buffers, 1) * 4 + 64 + 64)
        self.arena = SymmetricArena(total, self.device, tag="ps")
        A = self.arena
        # flag words (int64) inside the flags page
        self.F_WREADY, self.F_GREADY, self.F_ABORT, self.F_BUFS = 0, MAXB * 8, MAXB * 8 + MAXP * MAXB * 8, MAXB * 8 + MAXP * MAXB * 8 + 8
        self.flags_local = A.view(self.off_flags, 4096, torch.int64)
        # ---- device-side control block ------------------------------------------------------------------------
        self.ctl = torch.zeros(64, dtype=torch.int64, device=self.device)     # [0]=worker step, [1]=ps step, [2]=err, [8..]=counters (int32 pairs)
        self.ctl[0] = 1
        self.ctl[1] = 1
        self.p_wstep = self.ctl.data_ptr()
        self.p_psstep = self.ctl.data_ptr() + 8
        self.p_err = self.ctl.data_ptr() + 16
        self.p_counters = self.ctl.data_ptr() + 64           # int32 counters: [0..MAXB) reduce, [MAXB..2MAXB) push, [2MAXB] bcast
        self.p_sel = self.ctl.data_ptr() + 256               # [MAXB] int64: the grid-wide arrival decision per bucket (reduce_update_kernel)
        self._host_wstep = 0
        self._host_psstep = 0
        self._published = 0
        self._hp_dev = None
        self._hp_key = None
        self._side = torch.cuda.Stream(device=self.device)
        self._push_stream = torch.cuda.Stream(device=self.device)     # bucket pushes overlap the remaining backward
        self._push_forked = False
        self._pushed = [False] * self.NB
        self.zeroes_grads = topology == "ps"
        # K < P / kill / timeout: a worker can be left behind by the server and must FOLLOW its step (the reference's workers receive the step from the
        # PS every iteration, distributed_worker.py:127-137).  worker_fetch then blocks the host until the weights are there and returns the exact step
        # the device adopted; with K == P nobody can fall behind and the host just counts (fully asynchronous).
        self.follow_server = False
        self.zero_on_push = True              # push_kernel leaves zeros behind (accumulate-in-place modes); the worker clears it in gather mode
        # --comm-type Async: the PS only publishes a version; every worker PULLS the weights from the PS replica with P2P loads
        # (the reference's point-to-point weight fetch, distributed_worker.py:201-219, which has no sender in that snapshot)
        self.pull_weights = bool(pull_weights) and topology == "ps" and world > 1
        # Central PS in PULL mode (PSB200_PS_PULL=1): workers do not push; they only raise "gradients ready" on the server, and the server's
        # reduce kernel reads the gradient arenas of all workers itself — with the NVLS multicast mapping as ONE in-switch-reduced stream
        # (multimem.ld_reduce) instead of W-1 incast streams into its slots.  Needs bf16 gradients, every rank a worker (co-located) and K = P.
        self.pull_grads = (os.environ.get("PSB200_PS_PULL", "0") == "1" and topology == "ps" and world > 1 and self.wire == 0
                           and colocate and self.NB == 1)
        if self.pull_grads:
            self.zeroes_grads = False
        torch.cuda.synchronize(self.device)

    # ---- storage handed to the roles -------------------------------------------------------------------------------
    def worker_weight_storage(self, dtype):
        assert dtype == torch.bfloat16, "NVLink tier: worker replicas are bf16 (use --dtype bf16)"
        return self.arena.view(self.off_w, self.T * 2, torch.bfloat16)

    ps_weight_storage = lambda self: self.arena.view(self.off_w, self.T * 2, torch.bfloat16)   # noqa: E731

    def ps_slot_storage(self):
        """The P gradient slots the push kernel writes on the PS, as one strided ``[P, total]`` tensor in the wire dtype (bf16 / fp32;
        uint8 payload for the fp8 wire): ``GradientAccumulator`` addresses "layer l from worker k" as views of THIS memory instead of
        owning a second P x total slab."""
        if self.topology != "ps":
            return None
        dt = {0: torch.bfloat16, 1: torch.uint8, 2: torch.float32}[self.wire]
        es = {0: 2, 1: 1, 2: 4}[self.wire]
        flat = self.arena.view(self.off_slots, self.num_workers * self.slot_stride, dt)
        return flat.as_strided((self.num_workers, self.T), (self.slot_stride // es, 1))

    def worker_grad_storage(self, dtype):
        nb = self.T * (2 if dtype == torch.bfloat16 else 4)
        return self.arena.view(self.off_grads, nb, dtype)

    def _wptr(self, rank):
        return self.arena.peer_ptrs[rank] + self.off_w

    def _flag(self, rank, word_off):
        return self.arena.peer_ptrs[rank] + self.off_flags + word_off

    def sync_start_step(self, step=None):
        """Collective (every rank, once, before the first step): rank 0 announces the step it starts from (1, or last+1 after
        ``--resume``) and every rank seeds its device step counters with it — the flag protocol compares flag values with these
        counters, so a worker that restarted at 1 against a PS resumed at last+1 would wait forever."""
        start = [int(step) if step is not None else 1]
        if dist.is_initialized() and self.world_size > 1:
            dist.broadcast_object_list(start, src=0)
        st = int(start[0])
        if st != 1:
            torch.cuda.synchronize(self.device)
            self.ctl[0] = st
            self.ctl[1] = st
            self._host_wstep = st - 1
            self._host_psstep = st - 1
            torch.cuda.synchronize(self.device)
        return st

    # ---- PS side ---------------------------------------------------------------------------------------------------------
    def _dests(self):
        return list(self.worker_ranks)

    def _weight_dests(self):
        """Replicas the PS kernel stores into: everyone (push/multicast), or only its own when workers pull."""
        return [0] if self.pull_weights else list(self.worker_ranks)

    def ps_publish(self, step, weights):
        """Step 1 / resume: copy the PS's bf16 weights (already in its W region) to every replica and release the
        ready flags.  Later steps are published by the reduce-update kernel itself."""
        if self._published >= step:
            return
        if self.topology == "sharded" and self.rank != 0:
            self._published = step         # rank 0 seeds every replica (identical init everywhere)
            return
        assert weights.flat.data_ptr() == self.arena.local_ptr + self.off_w, "PS low-precision copy must live in the arena"
        d = self._dests()
        use_mc = self.arena.mc_ptr and len(d) == self.world_size and not self.pull_weights
        nfl = self.NB if self.topology == "ps" else self.world_size
        if self.pull_weights:      # nothing to copy (the PS replica is the source of truth); just release the version flags
            self.ext.bcast(self.arena.local_ptr + self.off_w, [self._wptr(r) for r in d], 0, 0, [self._flag(r, self.F_WREADY) for r in d], nfl,
                           self.p_psstep, 0, self.p_counters + 4 * 2 * MAXB)
        else:
            self.ext.bcast(self.arena.local_ptr + self.off_w, [self._wptr(r) for r in d], (self.arena.mc_ptr + self.off_w) if use_mc else 0,
                           self.T * 2, [self._flag(r, self.F_WREADY) for r in d], nfl, self.p_psstep, 0, self.p_counters + 4 * 2 * MAXB)
        ops.count_launch()
        self._published = step
        self.stats["bcast_bytes"] += self.T * 2

    def _hp(self, optimizer):
        g = optimizer.param_groups[0]
        fs = optimizer.flat_state()
        kind = 1 if "betas" in g else 0
        key = (kind, g['lr'], g.get('momentum', 0), g.get('dampening', 0), g['weight_decay'], g.get('nesterov', False),
               g.get('betas', (0, 0)), g.get('eps', 0), g.get('amsgrad', False))
        if self._hp_dev is None or key != self._hp_key:
            steps = fs["steps"]
            blob = self.ext.make_hp(g['lr'], g.get('momentum', 0.0), g.get('dampening', 0.0), g['weight_decay'], bool(g.get('nesterov', False)),
                                    g.get('betas', (0.9, 0.999))[0], g.get('betas', (0.9, 0.999))[1], g.get('eps', 1e-8), bool(g.get('amsgrad', False)), kind, steps)
            if self._hp_dev is None:
                self._hp_dev = blob.to(self.device)
            else:
                torch.cuda.current_stream().synchronize()
                self._hp_dev.copy_(blob)
            self._hp_key = key
        return self._hp_dev.data_ptr(), kind, fs

    def ps_reduce_update(self, step, k, optimizer, master, strict_reference=False, kill_stragglers=False, timeout_s=0.0, lowp=None):
        hp_ptr, kind, fs = self._hp(optimizer)
        P = self.num_workers
        if self.topology == "ps" and self.pull_grads:
            src = [self.arena.peer_ptrs[r] + self.off_grads for r in self.worker_ranks]
            scl, mode = [], (2 if getattr(self, "_grad_is_fp32", False) else 0)
            lo, hi = [0], [self.T]
            wready = [self._flag(r, self.F_WREADY) for r in self._dests()]
        elif self.topology == "ps":
            src = [self.arena.local_ptr + self.off_slots + s * self.slot_stride for s in range(P)]
            scl = [p + self.slot_q_bytes for p in src] if self.wire == 1 else []
            lo = [b[0] for b in self.buckets]
            hi = [b[1] for b in self.buckets]
            wready = [self._flag(r, self.F_WREADY) for r in self._dests()]
            mode = self.wire
        else:
            gdt = 2 if self._grad_is_fp32 else 0
            src = [self.arena.peer_ptrs[r] + self.off_grads for r in self.worker_ranks]
            scl, mode = [], gdt
            lo, hi = [self.shards[self.rank][0]], [self.shards[self.rank][1]]
            wready = [self._flag(r, self.F_WREADY + 8 * self.rank) for r in self._dests()]
        take = min(P, (k + 1) if strict_reference else k)
        d = self._dests()
        use_mc = bool(self.arena.mc_ptr) and len(d) == self.world_size and not self.pull_weights
        if kind == 0:
            s0, s1, s2 = fs["buf"], None, None
        else:
            s0, s1, s2 = fs["m"], fs["v"], fs["vmax"]
        nvls = self._nvls_reduce_ok(mode, take, len(src))
        if nvls:
            self.ext.ru_set_mc_src(self.arena.mc_ptr + self.off_grads)
        # --mode kill: the kernel raises the abort word of every worker whose gradient is not part of the decision (C10);
        # --mode timeout: after timeout_s it proceeds with the >= 1 gradients that arrived instead of failing the job
        abort_words = [self._flag(r, self.F_ABORT) for r in self.worker_ranks] if (kill_stragglers and self.topology == "ps") else []
        partial_ok = bool(timeout_s) and float(timeout_s) > 0.0
        try:
            self.ext.reduce_update(src, scl, int(take), 1.0 / float(k), int(mode), self.arena.local_ptr + self.off_flags + self.F_GREADY,
                                   self.p_psstep, lo, hi, master.flat, s0, s1, s2, hp_ptr, self._ru_dst(d),
                                   (self.arena.mc_ptr + self.off_w) if use_mc else 0, wready, self.p_counters, self.p_err,
                                   float(self.timeout_s), True, self.p_sel, partial_ok, abort_words, float(timeout_s or 0.0))
        finally:
            if nvls:
                self.ext.ru_set_mc_src(0)
        ops.count_launch()
        fs["steps"] += 1
        self._host_psstep += 1
        self._published = step + 1
        self.stats["bcast_bytes"] += self.T * 2

    def _nvls_reduce_ok(self, mode, take, nsrc) -> bool:
        """In pull mode (sharded topology, or the central PS with ``PSB200_PS_PULL=1``) the gradient arenas sit at the same symmetric offset on every
        rank, so the server can read ONE in-switch-reduced stream (``multimem.ld_reduce``) instead of one stream per peer.  Valid only when every
        device of the multicast group contributes: bf16 gradients, take == P == world size.  Validated on hardware in round 2
        (``profiles/run27_multi_gpu.md``); ``PSB200_NVLS_REDUCE=0`` selects the per-peer loads."""
        return (os.environ.get("PSB200_NVLS_REDUCE", "1") == "1"
                and (self.topology == "sharded" or getattr(self, "pull_grads", False)) and bool(self.arena.mc_ptr) and mode == 0
                and take == nsrc == self.world_size)

    def _ru_dst(self, d):
        # the kernel walks dst[] and w_ready[] with one index: in pull mode every entry stores to the PS replica (idempotent) while
        # each worker still gets its own version flag
        return [self._wptr(0 if self.pull_weights else r) for r in d]

    def ps_signal_kill(self, ranks, step):
        for r in ranks:
            self.ext.set_word(self._flag(r, self.F_ABORT), int(step))

    def ps_recv_buffers(self, step, numel):
        if numel == 0:
            return None
        t0 = time.time()
        while self._read_word(self.F_BUFS) < step:
            if time.time() - t0 > self.timeout_s:
                raise TimeoutError("BN statistics for step %d never arrived" % step)
            time.sleep(0.0005)
        return self.arena.view(self.off_bufs + 64, numel * 4, torch.float32).clone()

    # ---- worker side -----------------------------------------------------------------------------------------------------
    def fetch_wait_args(self):
        """What a consumer kernel needs to acquire this step's weights itself (``WorkerCompute.fetch_wait_args``): the weights-ready words of this
        rank, how many, the worker step word, the error word, the timeout.  Not offered in pull mode (the pull copy is the first reader)."""
        if self.pull_weights and self.rank != 0:
            return None
        n = self.NB if self.topology == "ps" else self.world_size
        return (self.arena.local_ptr + self.off_flags + self.F_WREADY, n, self.p_wstep, self.p_err, float(self.timeout_s))

    def worker_fetch(self, weights, timeout_s=0.0, fused_wait=False):
        """Enqueue the device-side wait for this step's weights.  Default: the host does not block and counts the steps itself.  ``fused_wait``: the
        first kernel of the caller's step acquires the ready words itself (ops/csrc/conv2_tcgen05.cu: weight_transpose_batched_kernel) — no wait kernel
        is launched here.  ``follow_server`` (K < P, kill / timeout modes): blocks until the weights are there and returns the step the device adopted
        (wait_flag_kernel moves a worker that fell behind to the server's step); -1 once the server has shut the job down."""
        n = self.NB if self.topology == "ps" else self.world_size
        pulls = self.pull_weights and self.rank != 0
        if self.follow_server or pulls or not fused_wait:
            self.ext.wait_flag(self.arena.local_ptr + self.off_flags + self.F_WREADY, n, self.p_wstep, self.p_err, self.timeout_s)
            ops.count_launch()
        else:
            self.stats["fused_waits"] = self.stats.get("fused_waits", 0) + 1
        if pulls:
            # P2P pull of the fresh weights from the PS replica into ours (no flags to release: nflags = 0)
            self.ext.bcast(self.arena.peer_ptrs[0] + self.off_w, [self.arena.local_ptr + self.off_w], 0, self.T * 2,
                           [self._flag(self.rank, self.F_WREADY)], 0, self.p_wstep, 0, self.p_counters + 4 * (2 * MAXB + 2))
            ops.count_launch()
        self._pushed = [False] * self.NB
        if self.follow_server:
            d = int(self.ctl[0].item())                # host sync: the wait kernel has finished, ctl[0] is the step this worker now computes
            self.check_error()
            if d >= DONE_STEP:
                return -1
            self._host_wstep = d
            return d
        self._host_wstep += 1
        return self._host_wstep

    def next_worker_step(self) -> int:
        """The step the next worker_fetch will at least ask for (exact in follow_server mode, where the host tracks the device counter)."""
        return self._host_wstep + 1

    def ps_shutdown(self, step=None):
        """End of training (C1's step = -1 in the message protocol): publish DONE_STEP as the weights-ready value on every worker.  Kernels still
        waiting for a publication pass at once, followers see it in their next worker_fetch and stop."""
        if self.rank != 0 and self.topology != "sharded":
            return
        d = self._dests()
        nfl = self.NB if self.topology == "ps" else self.world_size
        for r in d:
            for b in range(nfl):
                self.ext.set_word(self._flag(r, self.F_WREADY + 8 * b), DONE_STEP)
        ops.count_launch(len(d) * nfl)
        torch.cuda.synchronize(self.device)

    def _my_slot(self):
        return self.worker_slot(self.rank)

    def worker_push_bucket(self, b: int, grads: FlatBuffer, last_bucket: bool = True):
        lo, hi = self.buckets[b]
        slot = self._my_slot()
        fp32 = grads.dtype == torch.float32
        if self.topology == "ps" and self.pull_grads:
            self._grad_is_fp32 = fp32
            assert grads.flat.data_ptr() == self.arena.local_ptr + self.off_grads, "pull mode: gradients must live in the arena"
            # no data moves here: one release of "gradients of step s ready" in the server's flag page
            self.ext.bcast(self.arena.local_ptr + self.off_grads, [self.arena.peer_ptrs[0] + self.off_grads], 0, 0,
                           [self._flag(0, self.F_GREADY + 8 * (slot * MAXB))], 1, self.p_wstep, 0, self.p_counters + 4 * (2 * MAXB + 1))
            ops.count_launch()
        elif self.topology == "ps":
            base = self.arena.peer_ptrs[0] + self.off_slots + slot * self.slot_stride
            side = self.NB > 1 and not last_bucket
            if side:
                # fork: the push of a finished bucket runs on a side stream while backward continues on the main one
                # (inside a CUDA-graph capture this becomes a fork/join in the graph)
                ev = torch.cuda.Event()
                ev.record()
                self._push_stream.wait_event(ev)
                self._push_forked = True
            with torch.cuda.stream(self._push_stream if side else torch.cuda.current_stream()):
                self.ext.push(grads.flat.data_ptr(), fp32, base, base + self.slot_q_bytes, lo, hi,
                              self._flag(0, self.F_GREADY + 8 * (slot * MAXB + b)), self.p_wstep, self.p_counters + 4 * (MAXB + b), self.wire, self.zero_on_push)
            ops.count_launch()
            wbytes = {0: 2, 1: 1.03125, 2: 4}[self.wire]
            self.stats["push_bytes"] += int((hi - lo) * wbytes)
        else:
            self._grad_is_fp32 = fp32
            assert grads.flat.data_ptr() == self.arena.local_ptr + self.off_grads, "sharded topology: gradients must live in the arena"
            self._raise_ready(slot)      # pull mode: no data moves here, only "gradients ready" on every shard owner
        self._pushed[b] = True

    def _raise_ready(self, slot):
        # one tiny kernel per peer would be W launches; bcast_kernel with zero payload releases all flags in one launch
        d = self.worker_ranks
        self.ext.bcast(self.arena.local_ptr + self.off_grads, [self.arena.peer_ptrs[r] + self.off_grads for r in d], 0, 0,
                       [self._flag(r, self.F_GREADY + 8 * (slot * MAXB)) for r in d], 1, self.p_wstep, 0, self.p_counters + 4 * (2 * MAXB + 1))
        ops.count_launch()
        self.stats["push_bytes"] += 0

    def worker_push(self, step, grads, lo=0, hi=None, last=True):
        if hi is not None and not last:
            for b, (blo, bhi) in enumerate(self.buckets):
                if blo >= lo and bhi <= hi and not self._pushed[b]:
                    self.worker_push_bucket(b, grads, last_bucket=False)
            return
        for b in range(self.NB - 1, -1, -1):
            if not self._pushed[b]:
                self.worker_push_bucket(b, grads, last_bucket=True)
        if self._push_forked:                      # join the side stream before the step counter moves on
            ev = torch.cuda.Event()
            ev.record(self._push_stream)
            torch.cuda.current_stream().wait_event(ev)
            self._push_forked = False
        self.ext.add_step(self.p_wstep, 1)
        ops.count_launch()

    def join_push_stream(self) -> None:
        """Make the current stream wait for bucket pushes issued on the side push stream (called at the end of backward: inside a CUDA-graph
        capture every forked stream has to be joined before the capture ends)."""
        if self._push_forked:
            ev = torch.cuda.Event()
            ev.record(self._push_stream)
            torch.cuda.current_stream().wait_event(ev)
            self._push_forked = False

    def mark_pushed(self, b: int) -> None:
        """Bucket ``b`` was pushed by kernels inside a replayed CUDA graph (the Python hook that normally records it does not run)."""
        self._pushed[b] = True

    def worker_push_killed(self, step, grads: FlatBuffer = None):
        """The worker abandoned this step (kill signal / own deadline).  Gradients accumulate in place and are normally zeroed by the push
        kernel, so whatever backward produced before it was cut short must be cleared here or it would leak into the next step."""
        if self._push_forked:
            ev = torch.cuda.Event()
            ev.record(self._push_stream)
            torch.cuda.current_stream().wait_event(ev)
            self._push_forked = False
        if grads is not None:
            grads.flat.zero_()
        self.ext.add_step(self.p_wstep, 1)
        ops.count_launch()

    def _read_word(self, word_off) -> int:
        with torch.cuda.stream(self._side):
            v = self.flags_local[word_off // 8].item()
        return int(v)

    def worker_poll_kill(self, step):
        # the abort word holds the LAST step this worker was left out of: anything >= the step being computed means "the server has moved on"
        return self._read_word(self.F_ABORT) >= step

    def worker_send_buffers(self, step, flat):
        """BN running statistics -> the PS (checkpoint steps).  One launch of the copy kernel over raw peer pointers: it stores the payload into the
        PS's buffer region and then releases F_BUFS = step (a torch tensor cannot wrap peer-mapped memory: from_blob checks the owning device)."""
        if flat.numel() == 0:
            return
        assert flat.numel() <= self.n_buffers, "BN statistics (%d floats) exceed the %d reserved in the arena (pass n_buffers=)" % (flat.numel(), self.n_buffers)
        n = (flat.numel() + 3) // 4 * 4
        stage = getattr(self, "_buf_stage", None)
        if stage is None or stage.numel() < n:
            stage = self._buf_stage = torch.zeros(n, dtype=torch.float32, device=self.device)
        stage[: flat.numel()].copy_(flat.float())
        # worker_push already advanced the step word: flag value = *p_wstep - 1 = step
        self.ext.bcast(stage.data_ptr(), [self.arena.peer_ptrs[0] + self.off_bufs + 64], 0, n * 4, [self._flag(0, self.F_BUFS)], 1,
                       self.p_wstep, -1, self.p_counters + 4 * (2 * MAXB + 3))
        ops.count_launch()

    def gather_master(self, master: FlatBuffer) -> None:
        """Sharded topology: every rank holds the authoritative fp32 master only for its shard.  Collective (all ranks):
        afterwards every rank's ``master`` is complete (used by the checkpoint writer; rare, goes through gloo)."""
        if self.topology != "sharded" or self.world_size == 1:
            return
        torch.cuda.synchronize(self.device)
        for r, (lo, hi) in enumerate(self.shards):
            t = master.flat[lo:hi].detach().cpu() if r == self.rank else torch.empty(hi - lo, dtype=torch.float32)
            dist.broadcast(t, src=r)
            if r != self.rank:
                master.flat[lo:hi].copy_(t)

    # ---- all-reduce (DDP) ---------------------------------------------------------------------------------------------------
    def all_reduce_grads(self, grads: FlatBuffer):
        """In-place mean of the gradient arenas over all ranks: raise "ready" on every peer, then each rank reduces its
        shard from all peers (P2P loads) and stores the mean back into every peer's arena (kind = 2: no optimizer)."""
        assert self.topology == "sharded" and grads.dtype == torch.bfloat16
        assert grads.flat.data_ptr() == self.arena.local_ptr + self.off_grads
        if getattr(self, "_ar_hp", None) is None:
            self._ar_hp = self.ext.make_hp(0.0, 0.0, 0.0, 0.0, False, 0.0, 0.0, 0.0, False, 2, 0).to(self.device)
            self._ar_w = torch.zeros(8, dtype=torch.float32, device=self.device)     # unused master pointer
        self.worker_fetch(None)                         # previous round's stores into our arena have landed
        self._raise_ready(self._my_slot())
        d = self.worker_ranks
        lo, hi = self.shards[self.rank]
        nvls = self._nvls_reduce_ok(0, len(d), len(d))
        if nvls:
            self.ext.ru_set_mc_src(self.arena.mc_ptr + self.off_grads)
        try:
            self.ext.reduce_update([self.arena.peer_ptrs[r] + self.off_grads for r in d], [], len(d), 1.0 / len(d), 0,
                                   self.arena.local_ptr + self.off_flags + self.F_GREADY, self.p_psstep, [lo], [hi], self._ar_w, None, None, None,
                                   self._ar_hp.data_ptr(), [self.arena.peer_ptrs[r] + self.off_grads for r in d], 0,
                                   [self._flag(r, self.F_WREADY + 8 * self.rank) for r in d], self.p_counters, self.p_err, self.timeout_s, True,
                                   self.p_sel, False, [], 0.0)
        finally:
            if nvls:
                self.ext.ru_set_mc_src(0)
        ops.count_launch()
        self.ext.add_step(self.p_wstep, 1)
        # results are complete for the caller once every owner has published its shard
        self.ext.wait_flag(self.arena.local_ptr + self.off_flags + self.F_WREADY, self.world_size, self.p_wstep, self.p_err, self.timeout_s)

    # ---- misc --------------------------------------------------------------------------------------------------------------
    def check_error(self):
        e = int(self.ctl[2].item())
        if e:
            raise RuntimeError("NVLink transport: device-side wait timed out (code %d) — a peer is missing or dead" % e)

    def barrier(self):
        if dist.is_initialized():
            torch.cuda.synchronize(self.device)
            dist.barrier()

    def close(self):
        torch.cuda.synchronize(self.device)
        self.check_error()
