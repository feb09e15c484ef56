"""Process runtime: rank/role assignment, process-group bootstrap, transport selection, seeding.

Replaces ``mpirun`` + ``MPI.COMM_WORLD`` (``/root/reference/src/distributed_nn.py:70-126``): one process per GPU
(or per CPU rank on the gloo tier), rank 0 = parameter server, ranks 1..P = workers.  Ranks come from the
torchrun-style environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT) set by
``ps_pytorch_b200.launch`` or ``torch.distributed.run``.  torch.distributed is used only for rendezvous and
the CPU tier; on GPUs the gradient/weight traffic goes through the peer-memory kernels (``nvl.py``).
"""
from __future__ import annotations

import datetime
import os
import random

import numpy as np
import torch
import torch.distributed as dist

from ..config import num_classes_for
from ..models import build_model
from .arena import ParamLayout
from .compute import resolve_dtype
from .transport import GlooTransport, LocalTransport
from ..utils.data import prepare_data
from ..utils.obs import logger


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def seed_everything(seed: int, rank: int = 0) -> None:
    random.seed(seed + rank)
    np.random.seed(seed + rank)
    torch.manual_seed(seed)          # identical init on every rank; PS's init is the global init anyway


def parse_stragglers(spec: str):
    out = {}
    for item in filter(None, (spec or "").split(",")):
        r, ms = item.split(":")
        out[int(r)] = float(ms)
    return out


def init_process_group(use_cuda: bool, timeout_s: int = 600):
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if use_cuda:
            torch.cuda.set_device(local_rank)
        # gloo always: bootstrap/handle exchange + CPU tier.  NCCL is not used by the product path.
        dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=timeout_s))
    return rank, local_rank, world


def make_transport(args, layout: ParamLayout, device: torch.device, world: int, n_buffers: int = 0):
    kind = args.transport
    if kind == "auto":
        kind = "local" if world == 1 else ("nvl" if device.type == "cuda" else "gloo")
    compress = str(args.compress_grad).lower() == "compress"
    if kind == "local":
        return LocalTransport(layout, device)
    if kind == "gloo":
        return GlooTransport(layout, compress=compress, colocate=args.colocate_worker)
    if kind == "nvl":
        from .nvl import NvlTransport
        gd = args.grad_dtype
        if gd == "auto":
            gd = "fp8" if compress else "bf16"
        return NvlTransport(layout, device, grad_dtype=gd, colocate=args.colocate_worker, num_buckets=(args.num_buckets if args.overlap else 1),
                            pull_weights=str(args.comm_type).lower() == "async", n_buffers=n_buffers)
    raise ValueError(kind)


def run_ddp(args):
    """``--parallelism ddp``: every rank holds a replica, gradients are averaged by the all-reduce kernel (``ddp.py``) and every rank
    applies the same fused flat optimizer step to its own fp32 master — the capability of the reference's vendored
    ``DistributedDataParallel`` (``/root/reference/src/data_parallel_dist/data_parallel_dist.py:29-267``) as a runnable mode."""
    import time

    from .. import nn as pnn
    from ..optim import build_optimizer
    from ..utils import checkpoint as ckpt
    from ..utils.obs import format_worker_line
    from .arena import FlatBuffer
    from .ddp import DistributedDataParallel

    use_cuda = bool(args.enable_gpu) and not args.no_cuda and torch.cuda.is_available()
    rank, local_rank, world = init_process_group(use_cuda)
    device = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
    seed_everything(args.seed, rank)
    if args.backend != "auto":
        pnn.set_backend(args.backend)
    num_classes = num_classes_for(args.dataset)
    dtype = resolve_dtype(args.dtype, device)
    model = build_model(args.network, num_classes).to(device=device, dtype=dtype)
    ddp = DistributedDataParallel(model, device)
    master = ddp.weights if dtype == torch.float32 else FlatBuffer(ddp.layout, torch.float32, device)
    if master is not ddp.weights:
        master.flat.copy_(ddp.weights.flat)
    opt = build_optimizer(args.optimizer, model.parameters(), args.lr, args.momentum, weight_decay=args.weight_decay,
                          nesterov=args.nesterov, dampening=args.dampening, amsgrad=args.amsgrad)
    opt.attach_flat(master)
    criterion = pnn.CrossEntropyLoss()
    train_loader, _ = prepare_data(args, rank=rank, pin=use_cuda, out_dtype=torch.float32, channels_last=use_cuda)
    step, losses = 1, []
    model.train()
    while step <= args.max_steps:
        t0 = time.time()
        x, y = train_loader.next_batch()
        x = x.to(device=device, dtype=dtype, non_blocking=True)
        y = y.to(device, non_blocking=True)
        ddp.zero_grad()
        loss = criterion(ddp(x), y)
        loss.backward()
        ddp.sync_gradients()
        opt.step(ddp.grads.flat, out_lowp=None if master is ddp.weights else ddp.weights.flat)
        lv = float(loss)
        losses.append(lv)
        if step % max(1, args.log_interval) == 0:
            seen = (step * args.batch_size) % len(train_loader.dataset)
            logger.info(format_worker_line(rank, step, step * args.batch_size // max(1, len(train_loader.dataset)), seen, len(train_loader.dataset),
                                           lv, time.time() - t0, 0.0, 0.0, 0.0, 0.0))
        if rank == 0 and args.eval_freq > 0 and step % args.eval_freq == 0:
            ckpt.save_model(model, args.train_dir, step, master=master, optimizer=opt)
        step += 1
    if dist.is_initialized():
        dist.barrier()
    return {"ddp": ddp, "losses": losses, "master": master}


def run_sharded(args):
    """``--topology sharded``: every rank is a worker and the server of 1/W of the parameter arena (the scalable form of the PS; the
    reference pins ``n_masters`` to 1, ``/root/reference/tools/pytorch_ec2.py:26``).  All ranks run the same loop, so it is driven
    through ``engine.PSJob`` rather than the master/worker role loops."""
    import time

    from ..engine import PSJob
    from ..utils.obs import format_worker_line

    use_cuda = bool(args.enable_gpu) and not args.no_cuda and torch.cuda.is_available()
    if not use_cuda:
        raise SystemExit("--topology sharded needs the NVLink tier (GPUs): shards are served by peer-memory kernels")
    rank, local_rank, world = init_process_group(True)
    device = torch.device("cuda", local_rank)
    seed_everything(args.seed, rank)
    gd = args.grad_dtype
    if gd == "auto":
        gd = "bf16"
    job = PSJob(network=args.network, dataset=args.dataset, batch_size=args.batch_size, device=device, dtype=args.dtype, lr=args.lr,
                momentum=args.momentum, optimizer=args.optimizer, weight_decay=args.weight_decay, nesterov=args.nesterov, transport="nvl",
                topology="sharded", grad_dtype=gd, num_aggregate=None, cuda_graph=args.cuda_graph, backend=args.backend,
                max_steps=args.max_steps, eval_freq=args.eval_freq, train_dir=args.train_dir, seed=args.seed)
    train_loader, _ = prepare_data(args, rank=rank, pin=True, out_dtype=torch.float32, channels_last=True)
    n = len(train_loader.dataset)
    for step in range(1, args.max_steps + 1):
        t0 = time.time()
        x, y = train_loader.next_batch()
        loss = job.step(x, y)
        if step % max(1, args.log_interval) == 0:
            logger.info(format_worker_line(rank, step, step * args.batch_size // max(1, n), (step * args.batch_size) % max(1, n), n, float(loss),
                                           time.time() - t0, 0.0, 0.0, 0.0, 0.0))
        if args.eval_freq > 0 and step % args.eval_freq == 0:
            if rank == 0:                                # rank 0's own replica supplies the BN statistics of the checkpoint
                job.comm.worker_send_buffers(step, job.worker.compute.buffers_flat())
            job.master.cur_step = step
            job.master._save_model()                     # collective: gathers the shards, rank 0 writes model_step_<step>
            job.master.cur_step = step + 1
    job.close()
    if dist.is_initialized():
        dist.barrier()
    return {"job": job}


def run_distributed(args):
    """Body of ``distributed_nn.py``: build data, roles and transport from flags, then run."""
    if getattr(args, "parallelism", "ps") == "ddp":
        return run_ddp(args)
    if getattr(args, "topology", "ps") == "sharded":
        return run_sharded(args)
    use_cuda = bool(args.enable_gpu) and not args.no_cuda and torch.cuda.is_available()
    rank, local_rank, world = init_process_group(use_cuda)
    device = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
    seed_everything(args.seed, rank)
    if args.backend != "auto":
        from .. import nn as pnn
        pnn.set_backend(args.backend)
    num_classes = num_classes_for(args.dataset)
    dtype = resolve_dtype(args.dtype, device)
    proto = build_model(args.network, num_classes)
    layout = ParamLayout.from_module(proto, channels_last=(device.type == "cuda"))
    comm = make_transport(args, layout, device, world, n_buffers=sum(b.numel() for b in proto.buffers()))
    if hasattr(comm, "follow_server") and (args.mode != "normal" or 0 < int(args.num_aggregate or 0) < comm.num_workers):
        comm.follow_server = True      # K < P / kill / timeout: workers can be left behind and must follow the server's step
    stragglers = parse_stragglers(args.inject_straggler)
    failures = {int(r): int(st) for r, st in (item.split(":") for item in filter(None, (getattr(args, "inject_failure", "") or "").split(",")))}
    common = dict(batch_size=args.batch_size, learning_rate=args.lr, max_epochs=args.epochs, momentum=args.momentum,
                  network=args.network, comm_method=args.comm_type, eval_freq=args.eval_freq, train_dir=args.train_dir,
                  max_steps=args.max_steps, compress_grad=args.compress_grad, device=device, mode=args.mode, dtype=args.dtype)
    kwargs_master = dict(common, num_aggregate=args.num_aggregate, timeout_threshold=args.kill_threshold,
                         optimizer=args.optimizer, weight_decay=args.weight_decay, nesterov=args.nesterov,
                         dampening=args.dampening, amsgrad=args.amsgrad, strict_reference=args.strict_reference,
                         worker_dtype=dtype, resume=args.resume)
    kwargs_worker = dict(common, kill_threshold=args.kill_threshold, overlap=args.overlap, num_buckets=args.num_buckets,
                         cuda_graph=args.cuda_graph, straggle_ms=stragglers.get(rank, 0.0), fail_at_step=failures.get(rank, 0), metrics_port=getattr(args, "metrics_port", 0), log_interval=args.log_interval,
                         jsonl=(args.jsonl + ".rank%d" % rank) if args.jsonl else "")
    from .ps import SyncReplicasMaster_NN
    from .worker import DistributedWorker
    is_worker = rank in comm.worker_ranks
    train_loader = test_loader = None
    if is_worker:
        train_loader, test_loader = prepare_data(args, rank=rank, pin=use_cuda, out_dtype=torch.float32,
                                                 channels_last=use_cuda)
    result = {}
    if rank == 0:
        master = SyncReplicasMaster_NN(comm=comm, **kwargs_master)
        master.build_model(num_classes=num_classes, network=proto)
        print("I am the master: the world size is {}, cur step: {}".format(master.world_size, master.cur_step))
        if is_worker:
            w = DistributedWorker(comm=comm, **kwargs_worker)
            w.build_model(num_classes=num_classes)
            w.attach_loader(train_loader)
            master._colocated_worker = w
            result["worker"] = w
        comm.sync_start_step(master.cur_step)                    # collective: 1, or last+1 after --resume
        master.start()
        print("Done sending messages to workers!")
        result["master"] = master
    else:
        w = DistributedWorker(comm=comm, **kwargs_worker)
        w.build_model(num_classes=num_classes)
        print("I am worker: {} in all {} workers, next step: {}".format(w.rank, comm.num_workers, w.next_step))
        comm.sync_start_step(None)
        w.train(train_loader=train_loader, test_loader=test_loader)
        print("Now the next step is: {}".format(w.next_step))
        result["worker"] = w
    comm.close()
    if dist.is_initialized():
        dist.barrier()
    return result
