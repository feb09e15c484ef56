"""The parameter server (rank 0).

``SyncReplicasMaster_NN`` keeps the reference's class/method surface
(``/root/reference/src/sync_replicas_master_nn.py:96-287``): ``build_model`` → ``start`` drives
publish step + weights → gather K-of-P gradients → average → ``optimizer.step(grads)`` → periodic checkpoint.
The mechanism behind each method is the transport's (gloo messages on CPU; on B200 one fused
reduce+optimizer+multicast kernel over peer-mapped arenas, ``nvl.py``), and the PS state lives on the PS *GPU* in
fp32 (master weights + optimizer moments) rather than in CPU numpy.
"""
from __future__ import annotations

import time
from typing import List, Optional

import numpy as np
import torch

from .arena import FlatBuffer, ParamLayout
from .compute import load_buffers_flat
from .transport import Transport
from ..models import build_model
from ..optim import build_optimizer
from ..utils import checkpoint as ckpt
from ..utils.obs import logger

STEP_START_ = 1


class GradientAccumulator(object):
    """Per-layer arrival counters + per-(layer, worker) staging views — the TF "ConditionalAccumulator"
    analogue of the reference (``sync_replicas_master_nn.py:49-93``).  With flat arenas the staging buffers are
    views into one ``[P, total]`` slab, so ``gradient_aggregator[l][k]`` still addresses "layer l from worker k".  On the NVLink tier
    ``storage`` is ``NvlTransport.ps_slot_storage()``: the very slots the workers' push kernels store into (wire dtype), nothing is copied."""

    def __init__(self, layout: ParamLayout, num_worker: int, mode: str = 'None', device="cpu", dtype=torch.float32,
                 storage: Optional[torch.Tensor] = None):
        self._mode = mode
        self.layout = layout
        self.num_worker = num_worker
        self.slab = storage if storage is not None else torch.zeros(num_worker, layout.total, dtype=dtype, device=device)
        self.slots: List[FlatBuffer] = [FlatBuffer(layout, self.slab.dtype, self.slab.device, storage=self.slab[k]) for k in range(num_worker)]
        self.gradient_aggregator = [[self.slots[k].views[l] for k in range(num_worker)] for l in range(len(layout.numels))]
        self.gradient_aggregate_counter = [0] * len(layout.numels)
        self.model_index_range = list(range(len(layout.numels)))

    def meset_everything(self):
        self._meset_grad_counter()
        self._meset_grad_aggregator()

    def _meset_grad_counter(self):
        self.gradient_aggregate_counter = [0 for _ in self.gradient_aggregate_counter]

    def _meset_grad_aggregator(self):
        if self._mode != 'compress':
            self.slab.zero_()


class SyncReplicasMaster_NN(object):
    def __init__(self, comm: Transport = None, **kwargs):
        self.comm = comm
        self.world_size = comm.world_size if comm is not None else 1
        self.cur_step = STEP_START_
        self.lr = kwargs['learning_rate']
        self.momentum = kwargs['momentum']
        self.network_config = kwargs['network']
        self.comm_type = kwargs.get('comm_method', 'Bcast')
        self._timeout_threshold = kwargs.get('timeout_threshold', 0.0)
        self._num_grad_to_collect = comm.num_workers if comm is not None else 1
        self._eval_freq = kwargs.get('eval_freq', 50)
        self._train_dir = kwargs.get('train_dir', 'output/models/')
        # the reference feeds --num-aggregate through the 'kill_threshold' kw (distributed_nn.py:88); both names work
        self._expected_grad_to_recv = int(kwargs.get('num_aggregate', kwargs.get('kill_threshold', self._num_grad_to_collect)))
        self._max_steps = kwargs.get('max_steps', 10000)
        self._compress_grad = kwargs.get('compress_grad', 'None')
        self._device = torch.device(kwargs.get('device', 'cpu'))
        self._mode = kwargs.get('mode', 'normal')
        self._optimizer_name = kwargs.get('optimizer', 'sgd')
        self._opt_kw = dict(weight_decay=kwargs.get('weight_decay', 0.0), nesterov=kwargs.get('nesterov', False),
                            dampening=kwargs.get('dampening', 0.0), amsgrad=kwargs.get('amsgrad', False))
        self._strict_reference = kwargs.get('strict_reference', False)
        self._lowp_dtype = kwargs.get('worker_dtype', torch.float32)
        self._resume = kwargs.get('resume', False)
        self._colocated_worker = None      # set by the launcher when --colocate-worker
        self._model_shapes = []
        self.step_times: List[float] = []

    # ------------------------------------------------------------------------------------------------------
    def build_model(self, num_classes=10, network=None, layout: Optional[ParamLayout] = None):
        self.network = network if network is not None else build_model(self.network_config, num_classes)
        self.network.to(self._device)
        cl = self._device.type == "cuda"
        if cl:
            self.network = self.network.to(memory_format=torch.channels_last)
        self.layout = layout or (self.comm.layout if self.comm is not None else ParamLayout.from_module(self.network, channels_last=cl))
        self.master = FlatBuffer(self.layout, torch.float32, self._device)
        self.master.bind_params(self.network)           # PS parameters ARE the fp32 master arena
        self.optimizer = build_optimizer(self._optimizer_name, self.network.parameters(), self.lr, self.momentum, **self._opt_kw)
        self.optimizer.attach_flat(self.master)
        lowp_storage = self.comm.ps_weight_storage() if hasattr(self.comm, "ps_weight_storage") else None
        if self._lowp_dtype == torch.float32 and lowp_storage is None:
            self.lowp = self.master
        else:
            self.lowp = FlatBuffer(self.layout, self._lowp_dtype, self._device, storage=lowp_storage)
            self.lowp.flat.copy_(self.master.flat)
        # On the NVLink tier the accumulator's staging buffers ARE the slots the push kernel writes (no second P x total slab on the PS GPU);
        # on the CPU tier it owns a host slab like the reference's numpy buffers.
        slab = self.comm.ps_slot_storage() if hasattr(self.comm, "ps_slot_storage") else None
        nslot = max(1, self._num_grad_to_collect)
        self.grad_accumulator = None
        if slab is not None:
            self.grad_accumulator = GradientAccumulator(self.layout, slab.size(0), self._compress_grad, device=self._device, storage=slab)
        elif self._device.type == "cpu" and self.layout.total * nslot * 4 < (1 << 31):
            self.grad_accumulator = GradientAccumulator(self.layout, nslot, self._compress_grad, device=self._device)
        self.init_model_shapes()
        self._n_buffers = sum(b.numel() for b in self.network.buffers())
        if self._resume:
            self._try_resume()
        return self

    def init_model_shapes(self):
        self._model_shapes = [tuple(s) for s in self.layout.shapes]

    def _try_resume(self):
        last = ckpt.latest_step(self._train_dir)
        if last is None:
            return
        ckpt.load_model_into(self.network, ckpt.model_path(self._train_dir, last))
        side = ckpt.load_sidecar(self._train_dir, last)
        if side.get("optimizer") is not None:
            self.optimizer.load_flat_state_dict(side["optimizer"])
        self.cur_step = last + 1
        if self.lowp is not self.master:
            self.lowp.flat.copy_(self.master.flat)
        logger.info("Master: resumed from step %d", last)

    # ------------------------------------------------------------------------------------------------------
    def start(self):
        k = max(1, min(self._expected_grad_to_recv, self._num_grad_to_collect)) if not self._strict_reference else self._expected_grad_to_recv
        kill = self._mode in ("kill", "timeout")
        timeout = float(self._timeout_threshold) if self._mode == "timeout" else 0.0
        first = self.cur_step
        for i in range(first, self._max_steps + 1):
            t0 = time.time()
            logger.info("Master node is entering step: {}".format(i))
            self.async_bcast_step()                                   # C1 + C2 (one call: step word rides with the weights)
            if self._colocated_worker is not None:
                self._colocated_worker.run_one_step()
            self.comm.ps_reduce_update(self.cur_step, k, self.optimizer, self.master, strict_reference=self._strict_reference,
                                       kill_stragglers=kill, timeout_s=timeout, **self._ru_kwargs())
            if self.lowp is not self.master and not getattr(self.comm, "fused_lowp", False):
                self.lowp.flat.copy_(self.master.flat)
            if self.cur_step % self._eval_freq == 0:
                self._save_model(file_path=self._generate_model_path())
            self.step_times.append(time.time() - t0)
            self.cur_step += 1
        self.comm.ps_shutdown(self.cur_step)

    def _ru_kwargs(self):
        return {"lowp": self.lowp} if getattr(self.comm, "fused_lowp", False) else {}

    def async_bcast_step(self):
        """Publish the step id and the weights for it (reference: ``async_bcast_step`` + ``async_bcast_layer_weights_bcast``,
        ``sync_replicas_master_nn.py:210-225``)."""
        self.comm.ps_publish(self.cur_step, self.lowp)

    async_bcast_layer_weights_bcast = async_bcast_step

    def aggregate_gradient(self, gradient, layer_idx):
        """Reference API (``:239-241``): accumulate one layer's gradient into the aggregate buffer."""
        if not hasattr(self, "_grad_aggregate_buffer"):
            self._grad_aggregate_buffer = FlatBuffer(self.layout, torch.float32, self._device)
        g = torch.as_tensor(np.asarray(gradient) if not torch.is_tensor(gradient) else gradient).to(self._device)
        self._grad_aggregate_buffer.views[layer_idx].add_(g.reshape(self._model_shapes[layer_idx]))

    def _model_update(self):
        """Reference API (``:204-208``): average the aggregate buffer by K and apply the optimizer."""
        self.optimizer.step(grads=self._grad_aggregate_buffer.flat, scale=1.0 / float(self._expected_grad_to_recv))

    def meset_grad_buffer(self):
        if hasattr(self, "_grad_aggregate_buffer"):
            self._grad_aggregate_buffer.zero_()

    def _generate_model_path(self):
        return ckpt.model_path(self._train_dir, self.cur_step)

    def _save_model(self, file_path=None):
        """Single designated writer: PS master weights + BN statistics fetched from the first worker."""
        if hasattr(self.comm, "gather_master"):
            self.comm.gather_master(self.master)              # sharded topology: collect the other owners' shards
            if self.comm.rank != 0:
                return                                        # single writer
        bufs = self.comm.ps_recv_buffers(self.cur_step, self._n_buffers) if self.comm is not None else None
        if bufs is not None and bufs.numel() == self._n_buffers and self._n_buffers:
            load_buffers_flat(self.network, bufs.to(self._device))
        ckpt.save_model(self.network, self._train_dir, self.cur_step, master=self.master, optimizer=self.optimizer)
