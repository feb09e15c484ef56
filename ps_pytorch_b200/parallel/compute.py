"""Worker-side compute core: model replica bound to flat weight/gradient arenas, one training step
(zero-grad → forward → loss → backward) optionally captured in a CUDA graph, with per-bucket hooks that let the
transport push gradient buckets while the rest of backward is still running (the capability of the reference's
Split models, ``/root/reference/src/model_ops/resnet_split.py:365-501``, for *any* model)."""
from __future__ import annotations

import time
import os
from typing import Callable, List, Optional

import torch
import torch.nn as tnn

from .arena import FlatBuffer, ParamLayout
from .. import nn as pnn
from .. import ops as _ops
from ..models.layerwise import LayerwiseBackwardMixin, StragglerKilled
from ..utils.obs import PhaseTimer
from ..utils.data import mark_h2d_issued


# conv weight gradients land directly in the gradient arena (no hand-over + gather copy); PSB200_GRAD_SINKS=0 restores the gather of everything
_GRAD_SINKS = os.environ.get("PSB200_GRAD_SINKS", "1") != "0"


def resolve_dtype(name: str, device) -> torch.dtype:
    if name == "auto":
        return torch.bfloat16 if torch.device(device).type == "cuda" else torch.float32
    return {"fp32": torch.float32, "bf16": torch.bfloat16}[name]


class WorkerCompute:
    def __init__(self, network: tnn.Module, device, dtype: torch.dtype, layout: Optional[ParamLayout] = None,
                 weight_storage: Optional[torch.Tensor] = None, grad_dtype: Optional[torch.dtype] = None,
                 channels_last: Optional[bool] = None, use_graph: bool = False, grad_storage: Optional[torch.Tensor] = None,
                 zero_grads_in_step: bool = True):
        self.device = torch.device(device)
        self.dtype = dtype
        if self.device.type == "cuda":
            torch.backends.cudnn.benchmark = True     # library-path convolutions: let cuDNN pick by measurement
        self.channels_last = (self.device.type == "cuda") if channels_last is None else channels_last
        self.network = network.to(self.device).to(dtype)
        if self.channels_last:
            self.network = self.network.to(memory_format=torch.channels_last)
        self.layout = layout or ParamLayout.from_module(self.network, channels_last=self.channels_last)
        self.weights = FlatBuffer(self.layout, dtype, self.device, storage=weight_storage)
        self.weights.bind_params(self.network, keep_arena_values=weight_storage is not None)
        self.grads = FlatBuffer(self.layout, grad_dtype or dtype, self.device, storage=grad_storage)
        self.zero_grads_in_step = zero_grads_in_step
        # Launch-count hygiene on GPU (a ResNet-18 step is ~300 launches; at ~2.5 us of drain+launch each that is 10 % of it):
        #  * gather mode: parameters start backward with .grad = None, autograd then just *hands over* each gradient tensor (no
        #    accumulate kernel), and ONE multi-tensor copy moves all of them into the flat push arena (62 add kernels -> 1).
        #    Bucketed overlap needs gradients to land in the arena during backward, so it keeps the accumulate-in-place mode.
        #  * the 20 per-layer `num_batches_tracked += 1` kernels become one foreach add.
        self.gather_grads = self.device.type == "cuda"
        self._bn_counters = [m.num_batches_tracked for m in self.network.modules()
                             if isinstance(m, pnn.BatchNorm2d) and m.num_batches_tracked is not None] if self.device.type == "cuda" else []
        if self._bn_counters:
            for m in self.network.modules():
                if isinstance(m, pnn.BatchNorm2d):
                    m.external_counter = True
        self.grads.bind_grads(self.network)
        self._params = list(self.network.parameters())
        self.criterion = pnn.CrossEntropyLoss()
        self.timer = PhaseTimer(self.device)
        self.use_graph = use_graph and self.device.type == "cuda"
        if self.device.type == "cuda" and _ops.available():
            from ..ops import conv as _conv
            _conv.enable_batched_dgrad_operands(True)
        self._sinked = self._unsinked = None
        self.fetch_wait_args = None            # (flags_ptr, n, step_ptr, err_ptr, timeout_s) of the transport: see will_fuse_wait
        self._last_fused_wait = self._graph_fused_wait = False
        self.unfused_wait = False              # True: the caller launches its own (event-bracketable) wait kernel as well; the fused one then passes at once
        self._graph = None
        self._static_x = self._static_y = self._static_loss = None
        self._bucket_hooks: List = []
        self._on_bucket: Optional[Callable[[int], None]] = None
        self._kill_poll: Optional[Callable[[], bool]] = None
        self.kill_deadline: float = 0.0
        self.after_backward: Optional[Callable[[], None]] = None
        self._warming_up = False               # graph warm-up passes: bucket hooks must not push (or poll) for real
        self._capturing = False
        self.captured_buckets: List[int] = []  # buckets whose push kernels were recorded inside the CUDA graph

    # ---- overlap: bucket-complete callbacks during backward ---------------------------------------------
    def enable_bucket_callbacks(self, buckets, on_bucket: Callable[[int], None]) -> None:
        """``on_bucket(b)`` fires (inside backward) once every parameter of bucket ``b`` has its gradient."""
        for h in self._bucket_hooks:
            h.remove()
        self._bucket_hooks, self._on_bucket = [], on_bucket
        self._buckets = buckets
        params = list(self.network.parameters())
        self._pending = [0] * len(buckets)
        owner = {}
        for b, (_, _, idx) in enumerate(buckets):
            for i in idx:
                owner[i] = b

        def make(i):
            def hook(_p):
                if self._warming_up:
                    return                      # off-graph warm-up: no pushes, no kill polls (they used to push a bogus step)
                if self._kill_poll is not None and not self._capturing and (self._kill_poll() or (self.kill_deadline and time.time() > self.kill_deadline)):
                    raise StragglerKilled()
                b = owner[i]
                self._pending[b] -= 1
                if self._pending[b] == 0 and self._on_bucket is not None:
                    if self.device.type == "cuda":
                        from ..ops import libconv as _lc
                        _lc.join(self.device)          # weight gradients of this bucket may still be in flight on the side stream
                    self._on_bucket(b)
                    if self._capturing:
                        self.captured_buckets.append(b)
            return hook
        for i, p in enumerate(params):
            self._bucket_hooks.append(p.register_post_accumulate_grad_hook(make(i)))

    def will_fuse_wait(self) -> bool:
        """True when the NEXT train_step acquires the transport's weights-ready words inside its first kernel, so the caller must not launch
        its own wait kernel (``worker_fetch(..., fused_wait=True)``)."""
        if self.fetch_wait_args is None or self.unfused_wait or self.device.type != "cuda" or not _ops.available():
            return False
        if self.use_graph:
            return self._graph is not None and self._graph_fused_wait
        from ..ops import conv as _conv
        return _conv.refresh_ready()

    def enable_kill_poll(self, poll: Callable[[], bool]) -> None:
        self._kill_poll = poll
        if not self._bucket_hooks:
            lo_hi = self.layout.buckets(1)
            self.enable_bucket_callbacks(lo_hi, lambda b: None)

    def _arm_buckets(self):
        if self._bucket_hooks:
            self._pending = [len(idx) for (_, _, idx) in self._buckets]

    # ---- one step ----------------------------------------------------------------------------------------
    def _step_body(self, x, y):
        gather = self.gather_grads and not self._bucket_hooks and self._layerwise_comm is None
        if gather:
            for p in self._params:
                p.grad = None
            if self.device.type == "cuda" and _ops.available() and _GRAD_SINKS:
                from ..ops import conv as _conv
                if self._sinked is None:           # conv weight gradients are written by the wgrad kernels straight into the push source
                    self._sinked = set(_conv.set_grad_sinks(self._params, self.grads.views))
                    self._unsinked = [i for i in range(len(self._params)) if i not in self._sinked]
                _conv.activate_grad_sinks(True)
        elif self.zero_grads_in_step:
            self.grads.flat.zero_()
        if not gather and self._params[0].grad is None:
            self.grads.bind_grads(self.network)
        self._arm_buckets()
        if self.device.type == "cuda" and _ops.available():
            from ..ops import conv as _conv
            # one launch: the transposed filters every dgrad of this step reads; it is also the first reader of this step's weights and carries
            # the transport's weights-ready wait when the worker asked for that (fetch_wait_args)
            self._last_fused_wait = _conv.refresh_dgrad_operands(self.fetch_wait_args)
            from ..ops import elementwise as _ew
            _ew.bump_dropout_seed(self.device)      # (only when a Dropout layer has run on this device) fresh masks on every graph replay
        if self._bn_counters and self.network.training:
            torch._foreach_add_(self._bn_counters, 1)
        with self.timer.phase("forward"):
            logits = self.network(x)
            loss = self.criterion(logits, y)
        with self.timer.phase("backward"):
            if isinstance(self.network, LayerwiseBackwardMixin) and self._layerwise_comm is not None:
                # reference call pattern (nn_ops.py:64-70): loss on detached logits, then the layer-wise walk
                lg = torch.autograd.grad(loss, self.network.output[-1])[0]
                mode = self._layerwise_mode
                if mode == "kill":
                    _, killed = self.network.backward_signal_kill(lg, self._layerwise_comm, None, self._cur_step)
                elif mode == "timeout":
                    _, killed = self.network.backward_timeout_kill(lg, self._layerwise_comm, None, self._cur_step)
                else:
                    self.network.backward_normal(lg, self._layerwise_comm, None, self._cur_step)
                    killed = False
                if killed:
                    raise StragglerKilled()
            else:
                loss.backward()
            if self.device.type == "cuda":
                from ..ops import libconv as _lc
                _lc.join(self.device)                  # side-stream weight gradients land before anyone reads .grad
            if self.after_backward is not None:
                self.after_backward()                  # e.g. join the transport's bucket-push stream (must happen inside a graph capture)
            if gather:
                # what is left to gather: BatchNorm / classifier parameters (and convolutions a kernel of ours did not handle); a sinked weight
                # whose layer fell back to the library this step has a .grad after all and is copied like the rest
                if self._sinked:
                    from ..ops import conv as _conv
                    _conv.activate_grad_sinks(False)
                    written = _conv.sinks_written()
                    extra = [i for i in self._sinked if self._params[i].grad is not None]
                    # a sinked weight that ALSO has a .grad was used twice (shared layer): its second gradient is added to the first
                    for i in [i for i in extra if self._params[i].data_ptr() in written]:
                        self.grads.views[i].add_(self._params[i].grad)
                    todo = self._unsinked + [i for i in extra if self._params[i].data_ptr() not in written]
                else:
                    todo = range(len(self._params))
                dst = [self.grads.views[i] for i in todo]
                if dst:
                    torch._foreach_copy_(dst, [self._params[i].grad if self._params[i].grad is not None else torch.zeros_like(self.grads.views[i])
                                               for i in todo])
        return loss.detach(), logits.detach()

    _layerwise_comm = None
    _layerwise_mode = "normal"
    _cur_step = 0

    def set_layerwise(self, communicator, mode: str = "normal") -> None:
        self._layerwise_comm, self._layerwise_mode = communicator, mode

    def prepare_input(self, x: torch.Tensor, y: torch.Tensor):
        host = x if (self.device.type == "cuda" and not x.is_cuda) else None
        x = x.to(self.device, non_blocking=True)
        if host is not None:
            mark_h2d_issued(host)
        if x.dtype != self.dtype:
            x = x.to(self.dtype)
        if self.channels_last and x.dim() == 4:
            x = x.contiguous(memory_format=torch.channels_last)
        return x, y.to(self.device, non_blocking=True)

    # ---- input staging: the next batch's H2D copy runs on a copy stream while the current step computes ---------------------
    def stage(self, x: torch.Tensor, y: torch.Tensor) -> None:
        """Hint: ``(x, y)`` (pinned host tensors) is the batch of the NEXT ``train_step`` call.  Its host→device copy is issued now
        on a dedicated copy stream (double-buffered staging), so it overlaps the step that is currently running; the next step then
        starts with a device-to-device hand-off instead of a 6 MB PCIe transfer on the critical path."""
        if self.device.type != "cuda":
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staging = [None, None]
            self._stage_i = 0
        i = self._stage_i = self._stage_i ^ 1
        buf = self._staging[i]
        if buf is None or buf[0].shape != x.shape or buf[0].dtype != x.dtype:
            bx = torch.empty(x.shape, dtype=x.dtype, device=self.device)
            if self.channels_last and x.dim() == 4:
                bx = bx.contiguous(memory_format=torch.channels_last)
            buf = self._staging[i] = (bx, torch.empty(y.shape, dtype=y.dtype, device=self.device), torch.cuda.Event())
        cs = self._copy_stream
        cs.wait_stream(torch.cuda.current_stream(self.device))      # the staging buffer's previous consumer is enqueued before us
        with torch.cuda.stream(cs):
            buf[0].copy_(x, non_blocking=True)
            buf[1].copy_(y, non_blocking=True)
            buf[2].record(cs)
            mark_h2d_issued(x, cs)
        self._staged = (x.data_ptr(), y.data_ptr(), i)

    def _take_staged(self, x, y):
        st = self._staged
        if st is not None and st[0] == x.data_ptr() and st[1] == y.data_ptr():
            self._staged = None
            bx, by, ev = self._staging[st[2]]
            torch.cuda.current_stream(self.device).wait_event(ev)
            return bx, by
        return x, y

    _copy_stream = None
    _staged = None

    def train_step(self, x: torch.Tensor, y: torch.Tensor, step: int = 0):
        """Returns ``(loss, logits)`` device tensors.  Gradients land in ``self.grads``."""
        try:
            return self._train_step(x, y, step)
        except BaseException:
            if self._sinked:                       # an abandoned step (StragglerKilled, an error) must not leave the gradient sinks armed
                from ..ops import conv as _conv
                _conv.activate_grad_sinks(False)
            raise

    def _train_step(self, x: torch.Tensor, y: torch.Tensor, step: int = 0):
        self._cur_step = step
        self.network.train()
        if self.device.type == "cuda":
            x, y = self._take_staged(x, y)
        if not self.use_graph:
            x, y = self.prepare_input(x, y)
            return self._step_body(x, y)
        if self._graph is None:
            self._capture(x, y)
        self._static_x.copy_(x, non_blocking=True)       # H2D from the loader's pinned buffer, or D2D from the staged copy
        self._static_y.copy_(y, non_blocking=True)
        if not x.is_cuda:
            mark_h2d_issued(x)
        self._graph.replay()
        _ops.count_launch(self.graph_launches)
        return self._static_loss, self._static_logits

    def _capture(self, x, y):
        xs, ys = self.prepare_input(x, y)
        self._static_x, self._static_y = xs.clone(memory_format=torch.preserve_format), ys.clone()
        # The step is captured from a HIGH-priority stream; the weight gradients are forked to the default-priority side stream
        # (ops/libconv.py:_side).  Graph kernel nodes inherit the priority of the stream they were captured from, so whenever a critical-path
        # kernel (dgrad / BN backward) and a weight gradient are both ready, the block scheduler serves the critical path first and the weight
        # gradients fill in behind the memory-bound BatchNorm kernels.  PSB200_STEP_PRIORITY=0 disables.
        import os as _os
        prio = -1 if _os.environ.get("PSB200_STEP_PRIORITY", "1") != "0" else 0
        s = torch.cuda.Stream(priority=prio)
        s.wait_stream(torch.cuda.current_stream())
        saved = [b.detach().clone() for b in self.network.buffers()]      # warm-up must not advance BN statistics
        self._warming_up = True
        try:
            with torch.cuda.stream(s):
                for _ in range(3):   # warm-up off-graph: cuDNN autotune, lazy inits
                    self._step_body(self._static_x, self._static_y)
        finally:
            self._warming_up = False
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.no_grad():
            for b, v in zip(self.network.buffers(), saved):
                b.copy_(v)
        self.grads.flat.zero_()                                          # ... nor leave warm-up gradients in the push source
        self.timer.resolve()
        self._graph = torch.cuda.CUDAGraph()
        self.timer = _NullTimer()
        n0 = _ops.launch_count
        self.captured_buckets = []
        self._capturing = True
        try:
            with torch.cuda.graph(self._graph, stream=s):
                self._static_loss, self._static_logits = self._step_body(self._static_x, self._static_y)
        finally:
            self._capturing = False
        self._graph_fused_wait = self._last_fused_wait     # the recorded first kernel waits for the weights itself on every replay
        self.graph_launches = _ops.launch_count - n0      # our kernels recorded in the graph, re-counted per replay

    @torch.no_grad()
    def evaluate(self, loader, max_batches: int = 0):
        """Average NLL + Prec@1/Prec@5 over a loader (reference ``distributed_worker.py:274-299``)."""
        from ..utils.metrics import accuracy
        self.network.eval()
        tot_loss, p1, p5, nb, n = 0.0, 0.0, 0.0, 0, 0
        for x, y in loader:
            x, y = self.prepare_input(x, y)
            out = self.network(x).float()
            tot_loss += torch.nn.functional.cross_entropy(out, y, reduction='sum').item()
            a1, a5 = accuracy(out, y, topk=(1, 5))
            p1 += a1.item(); p5 += a5.item(); nb += 1; n += y.numel()
            if max_batches and nb >= max_batches:
                break
        self.network.train()
        return tot_loss / max(n, 1), p1 / max(nb, 1), p5 / max(nb, 1)

    def buffers_flat(self) -> torch.Tensor:
        """BN running statistics (and counters) as one fp32 vector — exchanged with the PS at checkpoint steps."""
        bs = [b.detach().reshape(-1).float() for b in self.network.buffers()]
        return torch.cat(bs) if bs else torch.zeros(0, device=self.device)


class _NullTimer:
    import contextlib as _c

    @_c.contextmanager
    def phase(self, name):
        yield

    def resolve(self):
        return {}

    def discard(self):
        pass


def load_buffers_flat(network: tnn.Module, flat: torch.Tensor) -> None:
    off = 0
    with torch.no_grad():
        for b in network.buffers():
            n = b.numel()
            b.copy_(flat[off:off + n].reshape(b.shape).to(b.dtype))
            off += n
