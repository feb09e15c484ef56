"""Operative baseline: the *reference algorithm* re-expressed with NCCL + stock ``torch.nn``
(cuDNN/cuBLAS/ATen) on the same box.  Rank 0 is the parameter server (fp32 master + SGD-momentum, no forward/backward),
ranks 1..N-1 are workers; per step and per parameter tensor: ``dist.reduce(grad, dst=0)`` → PS averages and steps →
``dist.broadcast(weight, src=0)`` — the per-layer message structure of
``/root/reference/src/sync_replicas_master_nn.py:218-237`` / ``distributed_worker.py:254-272`` with NCCL in place of
mpi4py.  With N=1 the PS and the single worker share the GPU (no communication).  Nothing from ``ps_pytorch_b200``'s
kernels or engine is on this path: it is the thing the product has to beat, not part of it."""
from __future__ import annotations

import json
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


class BasicBlock(nn.Module):
    def __init__(self, i, o, s):
        super().__init__()
        self.conv1 = nn.Conv2d(i, o, 3, s, 1, bias=False); self.bn1 = nn.BatchNorm2d(o)
        self.conv2 = nn.Conv2d(o, o, 3, 1, 1, bias=False); self.bn2 = nn.BatchNorm2d(o)
        self.shortcut = nn.Sequential()
        if s != 1 or i != o:
            self.shortcut = nn.Sequential(nn.Conv2d(i, o, 1, s, bias=False), nn.BatchNorm2d(o))

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + self.shortcut(x))


class ResNet18(nn.Module):
    def __init__(self, nc=10):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False); self.bn1 = nn.BatchNorm2d(64)
        layers, i = [], 64
        for o, s in [(64, 1), (64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1)]:
            layers.append(BasicBlock(i, o, s)); i = o
        self.layers = nn.Sequential(*layers)
        self.linear = nn.Linear(512, nc)

    def forward(self, x):
        out = self.layers(F.relu(self.bn1(self.conv1(x))))
        return self.linear(F.avg_pool2d(out, 4).flatten(1))


def run(a, rank, local, world, use_graph=True):
    """Times the baseline; returns the result dict on rank 0 (None elsewhere).  The caller owns the default process group.
    ``use_graph``: capture the step (NCCL collectives included) in a CUDA graph like the product does; falls back to eager if capture fails."""
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = True
    nccl = dist.new_group(backend="nccl") if world > 1 else None
    torch.manual_seed(1)
    model = ResNet18().to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    params = list(model.parameters())
    is_ps = rank == 0
    is_worker = world == 1 or rank > 0
    nworkers = max(1, world - 1)
    master = [p.detach().float().clone() for p in params] if is_ps else None
    mom = [torch.zeros_like(m) for m in master] if is_ps else None
    B = a.batch_size
    x = torch.randn(B, 3, 32, 32, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (B,), device=dev)
    lr, mu = 0.1, 0.9
    zero_g = [torch.zeros_like(p) for p in params] if (world > 1 and not is_worker) else None

    def step(first):
        if is_worker:
            for p in params:
                p.grad = None
            loss = F.cross_entropy(model(x).float(), y)
            loss.backward()
        if world > 1:
            for i, p in enumerate(params):                    # per-layer push (reduce to the PS)
                if is_worker:
                    g = p.grad
                else:
                    g = zero_g[i].zero_()
                dist.reduce(g, dst=0, group=nccl)
                if is_ps:
                    p.grad = g
        if is_ps:
            for p, m, b in zip(params, master, mom):           # PS: average + SGD momentum on the fp32 master
                d = p.grad.float() / nworkers
                if first:
                    b.copy_(d)
                else:
                    b.mul_(mu).add_(d)
                m.add_(b, alpha=-lr)
                p.data.copy_(m)
        if world > 1:
            for p in params:                                  # per-layer weight broadcast
                dist.broadcast(p.data, src=0, group=nccl)

    for i in range(max(3, a.warmup)):
        step(i == 0)
    torch.cuda.synchronize()
    graph = None
    if use_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step(False)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step(False)
            graph = g
        except Exception:                      # noqa: BLE001  (e.g. NCCL capture unsupported): eager, as the reference would run
            graph = None
            torch.cuda.synchronize()
    ok = torch.tensor([1 if graph is not None else 0])
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # all ranks must take the same path
    if int(ok[0]) == 0:
        graph = None

    def one():
        if graph is not None:
            graph.replay()
        else:
            step(False)
    for _ in range(2):
        one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    ms = float(t[0])
    return {"impl": "nccl-baseline", "metric": "ResNet-18/CIFAR-10 images/sec (whole box, device-timed, max over ranks)",
            "value": nworkers * B * a.steps / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "dtype": "bf16",
            "data": "synthetic", "cuda_graph": graph is not None,
            "config": {"model": "ResNet18", "global_batch": nworkers * B,
                       "parallelism": "1 PS + %d workers, NCCL reduce/broadcast per layer, stock torch.nn" % nworkers}}


def bench(a, rank, local, world):
    out = run(a, rank, local, world, use_graph=not getattr(a, "no_graph", False))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0
