"""Gradient-push + parameter-broadcast sweep (BASELINE.json config #5): 1 KB – 1 GB per worker at 2/4/8 GPUs,
our fused kernels over peer memory vs the reference algorithm on NCCL (reduce to rank 0 + PS update + broadcast).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29500 scripts/bw_sweep.py

One cycle = every worker pushes `bytes` of bf16 gradients, the PS averages them and applies SGD-momentum to the fp32
master, every worker receives the new bf16 weights.  Device-timed (CUDA events), max over ranks, 3 warm-up + `iters` timed
cycles.  The PS ingress is the bound: (W-1) x bytes must enter rank 0 over NVLink (measured peer-copy peak 770 GB/s)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200.optim import build_optimizer  # noqa: E402
from ps_pytorch_b200.parallel.arena import FlatBuffer, ParamLayout  # noqa: E402
from ps_pytorch_b200.parallel.nvl import NvlTransport  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("gloo", rank=rank, world_size=world)
nccl = dist.new_group(backend="nccl")
LINK = 770e9
sizes = [1 << k for k in range(10, 31, 2)]            # bytes per worker (bf16 payload)
wires = os.environ.get("WIRES", "bf16,fp8,pull").split(",")
rows = []


def max_over_ranks(ms):
    t = torch.tensor([ms], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize(); dist.barrier()
    return max_over_ranks(a.elapsed_time(b) / iters)


for nbytes in sizes:
    T = max(128, nbytes // 2 // 128 * 128)
    iters = 50 if nbytes <= (1 << 24) else (10 if nbytes <= (1 << 28) else 4)
    row = {"bytes": T * 2, "world": world}
    lay = ParamLayout(["w"], [(T,)], [T], [0], [False], T)
    for wire in wires:
        # "pull": central PS in pull mode — workers only raise a flag, the PS kernel reads ONE in-switch-reduced stream (multimem.ld_reduce)
        os.environ["PSB200_PS_PULL"] = "1" if wire == "pull" else "0"
        tr = NvlTransport(lay, dev, grad_dtype="bf16" if wire == "pull" else wire, colocate=True)
        grads = FlatBuffer(lay, torch.bfloat16, dev, storage=tr.worker_grad_storage(torch.bfloat16))
        lowp = FlatBuffer(lay, torch.bfloat16, dev, storage=tr.worker_weight_storage(torch.bfloat16))
        master = opt = None
        if rank == 0:
            master = FlatBuffer(lay, torch.float32, dev)
            opt = build_optimizer("sgd", [torch.nn.Parameter(master.flat)], lr=0.01, momentum=0.9)
            opt.attach_flat(master)
        state = {"step": 0}

        def cycle():
            state["step"] += 1
            s = state["step"]
            if rank == 0:
                tr.ps_publish(s, lowp)
            tr.worker_fetch(lowp)
            tr.worker_push(s, grads)
            if rank == 0:
                tr.ps_reduce_update(s, world, opt, master, lowp=lowp)
        ms = timed(cycle, iters)
        tr.check_error()
        row["ours_%s_us" % wire] = ms * 1e3
        row["ours_%s_ps_ingress_GBs" % wire] = (world - 1) * T * 2 / (ms * 1e-3) / 1e9
        row["backend"] = tr.arena.backend
        arena = tr.arena
        del tr, grads, lowp
        arena.close()
        torch.cuda.empty_cache()
    # NCCL arm: reduce to rank 0, PS-side SGD-momentum on fp32 master, broadcast bf16 weights
    g = torch.zeros(T, dtype=torch.bfloat16, device=dev)
    w = torch.zeros(T, dtype=torch.bfloat16, device=dev)
    if rank == 0:
        m32, buf = torch.zeros(T, device=dev), torch.zeros(T, device=dev)

    def nccl_cycle():
        dist.reduce(g, dst=0, group=nccl)
        if rank == 0:
            buf.mul_(0.9).add_(g.float(), alpha=1.0 / world)
            m32.add_(buf, alpha=-0.01)
            w.copy_(m32)
        dist.broadcast(w, src=0, group=nccl)
    ms = timed(nccl_cycle, iters)
    row["nccl_us"] = ms * 1e3
    row["nccl_ps_ingress_GBs"] = (world - 1) * T * 2 / (ms * 1e-3) / 1e9
    row["ours_bf16_frac_of_link"] = row.get("ours_bf16_ps_ingress_GBs", 0) * 1e9 / LINK
    row["speedup_vs_nccl_bf16"] = row["nccl_us"] / row["ours_bf16_us"] if "ours_bf16_us" in row else None
    row["speedup_vs_nccl_pull"] = row["nccl_us"] / row["ours_pull_us"] if "ours_pull_us" in row else None
    rows.append(row)
    if rank == 0:
        print(json.dumps(row), flush=True)
    del g, w
    torch.cuda.empty_cache()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/bw_sweep_n%d.json" % world, "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
