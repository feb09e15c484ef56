"""Workers (ranks 1..P, plus optionally a co-located one on the PS GPU).

``DistributedWorker`` keeps the reference's surface (``/root/reference/src/distributed_worker.py:67-307``):
``build_model`` → ``train(train_loader, test_loader)``; per step: fetch step + weights, forward, backward, push
gradients, log one line with the reference's field names.  Differences: the model replica's parameters and
gradients are views of flat arenas, so "fetch weights" is the PS writing into this replica (no ``load_state_dict``
rebuild), and "send grads" is one push per bucket that can overlap the remaining backward (``--overlap``).
"""
from __future__ import annotations

import os
import time
import torch

from .compute import WorkerCompute, resolve_dtype
from .transport import Transport
from ..models import build_model
from ..models.layerwise import LayerwiseBackwardMixin, StragglerKilled
from ..utils import checkpoint as ckpt
from ..utils.metrics import accuracy  # noqa: F401  (the reference defines accuracy() in this module; kept importable from here)
from ..utils.obs import JsonlSink, MetricsExporter, format_worker_line, logger

# the step's first kernel acquires the weights-ready words itself instead of a separate wait kernel (PSB200_FUSE_WAIT=0: always the wait kernel)
_FUSE_WAIT = os.environ.get("PSB200_FUSE_WAIT", "1") != "0"

STEP_START_ = 1


class ModelBuffer(object):
    """Per-layer "step last updated" bookkeeping of the reference (``distributed_worker.py:40-64``).  The receive
    scratch itself is gone: weights arrive in place in the replica's arena, ``recv_buf[l]`` are views of it."""

    def __init__(self, network, weights=None):
        self.recv_buf = list(weights.views) if weights is not None else [p.data for p in network.parameters()]
        self.layer_cur_step = [0 for _ in self.recv_buf]
        self.layer_shape = [tuple(p.size()) for p in network.parameters()]


class _BucketPusher:
    """Communicator handed to layer-wise backward / bucket hooks: pushes finished parameter ranges."""

    def __init__(self, worker: "DistributedWorker"):
        self.w = worker

    def push_params(self, indices, step):
        lo, hi = self.w.compute.layout.index_range(indices)
        self.w.comm.worker_push(step, self.w.compute.grads, lo, hi, last=False)

    def poll_kill(self, step):
        return self.w.comm.worker_poll_kill(step)


class DistributedWorker(object):
    def __init__(self, comm: Transport, **kwargs):
        self.comm = comm
        self.world_size = comm.world_size
        self.rank = comm.rank
        self.cur_step = 0
        self.next_step = 0
        self.batch_size = kwargs['batch_size']
        self.max_epochs = kwargs['max_epochs']
        self.momentum = kwargs['momentum']
        self.lr = kwargs['learning_rate']
        self._max_steps = kwargs['max_steps']
        self.network_config = kwargs['network']
        self.comm_type = kwargs.get('comm_method', 'Bcast')
        self.kill_threshold = kwargs.get('kill_threshold', 7.0)
        self._eval_freq = kwargs.get('eval_freq', 50)
        self._train_dir = kwargs.get('train_dir', 'output/models/')
        self._compress_grad = kwargs.get('compress_grad', 'None')
        self._device = torch.device(kwargs.get('device', 'cpu'))
        self._mode = kwargs.get('mode', 'normal')
        self._dtype = resolve_dtype(kwargs.get('dtype', 'auto'), self._device)
        self._overlap = kwargs.get('overlap', False)
        self._num_buckets = kwargs.get('num_buckets', 4)
        self._use_graph = kwargs.get('cuda_graph', False) and self._mode == 'normal'
        self._straggle_ms = float(kwargs.get('straggle_ms', 0.0))
        self._fail_at_step = int(kwargs.get('fail_at_step', 0))           # fault injection (--inject-failure), first attempt only
        self._log_interval = max(1, int(kwargs.get('log_interval', 1)))
        self._jsonl = JsonlSink(kwargs.get('jsonl') or None)
        self._metrics = MetricsExporter(int(kwargs.get('metrics_port', 0) or 0), self.rank)
        self._is_ckpt_source = comm.worker_ranks[0] == comm.rank
        self.losses = []
        self.step_times = []

    # ------------------------------------------------------------------------------------------------------
    def build_model(self, num_classes=10, network=None):
        net = network if network is not None else build_model(self.network_config, num_classes)
        if getattr(net, "kill_threshold", None) is not None:
            net.kill_threshold = self.kill_threshold
        storage = self.comm.worker_weight_storage(self._dtype) if hasattr(self.comm, "worker_weight_storage") else None
        gstorage = self.comm.worker_grad_storage(self._dtype) if hasattr(self.comm, "worker_grad_storage") else None
        self.compute = WorkerCompute(net, self._device, self._dtype, layout=self.comm.layout, weight_storage=storage,
                                     use_graph=self._use_graph, grad_storage=gstorage,
                                     zero_grads_in_step=not getattr(self.comm, "zeroes_grads", False))
        # (not when the worker FOLLOWS the server's step — K < P, kill / timeout: the blocking fetch has already adopted a step and told the host; a
        # second wait that may adopt a newer publication behind the host's back would desynchronise them)
        if _FUSE_WAIT and hasattr(self.comm, "fetch_wait_args") and not getattr(self.comm, "follow_server", False):
            self.compute.fetch_wait_args = self.comm.fetch_wait_args()
        self.network = self.compute.network
        self.optimizer = torch.optim.SGD(self.network.parameters(), lr=self.lr, momentum=self.momentum)  # zero_grad only, as in the reference
        self.criterion = self.compute.criterion
        self.init_recv_buf()
        self._pusher = _BucketPusher(self)
        if isinstance(self.network, LayerwiseBackwardMixin) and (self._overlap or self._mode != 'normal'):
            self.compute.set_layerwise(self._pusher, self._mode)
        elif self._overlap:
            self._buckets = self.compute.layout.buckets(self._num_buckets)
            self.compute.enable_bucket_callbacks(
                self._buckets, lambda b: self.comm.worker_push(self.cur_step, self.compute.grads, self._buckets[b][0], self._buckets[b][1], last=False))
            self.compute.after_backward = getattr(self.comm, "join_push_stream", None)
        if self._mode in ('kill', 'timeout') and not isinstance(self.network, LayerwiseBackwardMixin):
            self.compute.enable_kill_poll(lambda: self.comm.worker_poll_kill(self.cur_step))
        c = self.compute
        if hasattr(self.comm, "zero_on_push") and c.gather_grads and not c._bucket_hooks and c._layerwise_comm is None:
            # gather mode overwrites every gradient element each step: the push kernel need not write zeros behind itself
            self.comm.zero_on_push = False
        return self

    def init_recv_buf(self):
        self.model_recv_buf = ModelBuffer(self.network, self.compute.weights)

    # ------------------------------------------------------------------------------------------------------
    def sync_fetch_step(self):
        self.next_step = self.comm.worker_fetch(self.compute.weights, fused_wait=self.compute.will_fuse_wait())

    async_fetch_step = sync_fetch_step

    def async_fetch_weights_bcast(self):
        """Weights already landed with the step word (``worker_fetch``); only the per-layer step bookkeeping remains."""
        for i in range(len(self.model_recv_buf.layer_cur_step)):
            self.model_recv_buf.layer_cur_step[i] = self.cur_step

    def update_step(self):
        changed = (self.cur_step != self.next_step)
        self.cur_step = self.next_step
        return changed

    # ------------------------------------------------------------------------------------------------------
    def train(self, train_loader, test_loader=None):
        logger.info("Worker {}: starting training".format(self.rank))
        total = len(train_loader.dataset)
        self._done = False
        for num_epoch in range(self.max_epochs):
            it = iter(train_loader)
            nxt = next(it, None)
            batch_idx = 0
            while nxt is not None:
                xb, yb = nxt
                nxt = next(it, None)
                ok = self._one_step(xb, yb, num_epoch, batch_idx, total, nxt)
                batch_idx += 1
                if not ok:
                    self._done = True
                    break
            if self._done:
                break
        self._jsonl.close()

    def _one_step(self, xb, yb, num_epoch=0, batch_idx=0, total=0, next_batch=None) -> bool:
        iter_start = time.time()
        t0 = time.time()
        self.sync_fetch_step()                                            # C1 + C2
        if self.next_step < 0 or self.next_step > self._max_steps:
            return False
        self.update_step()
        self.async_fetch_weights_bcast()
        fetch_dur = time.time() - t0
        if self._straggle_ms > 0:
            time.sleep(self._straggle_ms * 1e-3)                          # fault injection (--inject-straggler)
        if self._fail_at_step and self.cur_step == self._fail_at_step and os.environ.get("PSB200_RESTART_COUNT", "0") == "0":
            logger.error("Worker %d: injected failure at step %d", self.rank, self.cur_step)
            os._exit(17)
        if self._mode == 'timeout':
            # the worker's own deadline for this step; not on its first iteration, which pays for lazy kernel loading / autotuning (seconds): with every
            # worker abandoning step 1 the server would never see a gradient
            self.compute.kill_deadline = (time.time() + float(self.kill_threshold)) if getattr(self, "_iters_done", 0) > 0 else 0.0
        self._iters_done = getattr(self, "_iters_done", 0) + 1
        killed = False
        try:
            loss, logits = self.compute.train_step(xb, yb, self.cur_step)
            if next_batch is not None and next_batch[0].is_pinned():
                self.compute.stage(next_batch[0], next_batch[1])       # overlap the next batch's H2D with this step
        except StragglerKilled:
            killed = True
        c0 = time.time()
        if killed:
            self.comm.worker_push_killed(self.cur_step, self.compute.grads)
            logger.info("Worker %d: step %d abandoned on kill signal", self.rank, self.cur_step)
        else:
            for b in self.compute.captured_buckets:      # pushed by kernels inside the replayed CUDA graph, not by the Python hook
                self.comm.mark_pushed(b)
            self._send_grads()
        if self.cur_step % self._eval_freq == 0 and self._is_ckpt_source:
            self.comm.worker_send_buffers(self.cur_step, self.compute.buffers_flat())
        log_now = (self.cur_step % self._log_interval == 0) or self.cur_step == 1
        if log_now and not killed:
            ph = self.compute.timer.resolve()
            loss_v = float(loss)
            comm_dur = time.time() - c0
            dur = time.time() - iter_start
            self.losses.append(loss_v)
            self.step_times.append(dur)
            msg_mb = self.comm.stats["push_bytes"] / 1e6
            self.comm.stats["push_bytes"] = 0
            line = format_worker_line(self.rank, self.cur_step, num_epoch, batch_idx * self.batch_size, total, loss_v, dur,
                                      fetch_dur, ph.get("forward", 0.0), ph.get("backward", 0.0), comm_dur, msg_mb,
                                      self.batch_size / max(dur, 1e-9))
            logger.info(line)
            self._jsonl.write(rank=self.rank, step=self.cur_step, loss=loss_v, time=dur, fetch=fetch_dur,
                              forward=ph.get("forward", 0.0), backward=ph.get("backward", 0.0), comm=comm_dur, msg_mb=msg_mb)
            self._metrics.update(self.cur_step, loss_v, dur, self.batch_size / max(dur, 1e-9), ph.get("forward", 0.0), ph.get("backward", 0.0), comm_dur, msg_mb)
        elif not log_now:
            self.compute.timer.discard()
        return self.cur_step < self._max_steps

    def run_one_step(self):
        """Used when this worker is co-located with the PS: the PS loop drives it."""
        if not hasattr(self, "_colo_iter"):
            self._colo_iter = iter(self._colo_loader)
        try:
            xb, yb = next(self._colo_iter)
        except StopIteration:
            self._colo_iter = iter(self._colo_loader)
            xb, yb = next(self._colo_iter)
        self._one_step(xb, yb, 0, 0, len(self._colo_loader.dataset))

    def attach_loader(self, loader):
        self._colo_loader = loader

    def _send_grads(self):
        """C6: hand the gradient arena (or its last bucket) to the transport."""
        self.comm.worker_push(self.cur_step, self.compute.grads, 0, None, last=True)

    def _evaluate_model(self, test_loader):
        loss, p1, p5 = self.compute.evaluate(test_loader)
        print('Test set: Step: {}, Average loss: {:.4f}, Prec@1: {} Prec@5: {}'.format(self.cur_step, loss, p1, p5))
        return loss, p1, p5

    def _generate_model_path(self):
        return ckpt.model_path(self._train_dir, self.cur_step)

    def _save_model(self, file_path=None):
        ckpt.save_model(self.network, self._train_dir, self.cur_step)
