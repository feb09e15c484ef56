"""Flag-protocol litmus (run under the launcher on >= 2 GPUs): hundreds of push -> reduce-update -> broadcast cycles on a small
arena with integer-valued payloads, so every intermediate result is exactly representable and ANY stale read, torn write or
missed flag shows up as an exact mismatch.  Worker k pushes the constant (k+1)*step; with SGD lr=1, momentum 0 the master after
step t must equal  -sum_{s<=t} s*mean(k+1)  everywhere, and every replica must hold its bf16 rounding."""
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
topology = sys.argv[1] if len(sys.argv) > 1 else "ps"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
pull = len(sys.argv) > 3 and sys.argv[3] == "pull"
T = 128 * 257
lay = ParamLayout(["w"], [(T,)], [T], [0], [False], T)
tr = NvlTransport(lay, dev, grad_dtype="fp32", colocate=True, topology=topology, pull_weights=pull)
grads = FlatBuffer(lay, torch.float32, dev, storage=tr.worker_grad_storage(torch.float32))
lowp = FlatBuffer(lay, torch.bfloat16, dev, storage=tr.worker_weight_storage(torch.bfloat16))
master = opt = None
if rank == 0 or topology == "sharded":
    master = FlatBuffer(lay, torch.float32, dev)
    opt = build_optimizer("sgd", [torch.nn.Parameter(master.flat)], lr=1.0, momentum=0.0)
    opt.attach_flat(master)
mean_k = sum(range(1, world + 1)) / world
expect = 0.0
bad = 0
for s in range(1, steps + 1):
    if master is not None:
        tr.ps_publish(s, lowp)
    tr.worker_fetch(lowp)
    # the replica every worker computes with must be the result of step s-1 (bf16-rounded)
    if s % 7 == 0 or s == steps:
        torch.cuda.synchronize()
        want = torch.tensor(expect, dtype=torch.float32).to(torch.bfloat16)
        if not bool((lowp.flat == want.to(dev)).all()):
            bad += 1
    grads.flat.fill_(float((tr.worker_slot(rank) + 1) * (s % 5)))
    tr.worker_push(s, grads)
    if master is not None:
        tr.ps_reduce_update(s, world, opt, master, lowp=lowp)
    expect -= (s % 5) * mean_k
torch.cuda.synchronize()
tr.check_error()
if master is not None:
    lo, hi = (0, T) if topology == "ps" else tr.shards[rank]
    if not bool((master.flat[lo:hi] == expect).all()):
        bad += 100
t = torch.tensor([bad])
dist.all_reduce(t)
if rank == 0:
    print("LITMUS", topology, "pull" if pull else "push", "steps", steps, "mismatches", int(t[0]))
dist.barrier()
sys.exit(1 if int(t[0]) else 0)
