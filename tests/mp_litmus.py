"""Flag-protocol litmus (run under the launcher on >= 2 GPUs): hundreds of push -> reduce-update -> broadcast cycles with integer-valued
payloads, so every intermediate result is exactly representable and ANY stale read, torn write or missed flag shows up as an exact
mismatch.  Worker k pushes (k+1) * (step % 5) * pattern, where `pattern` (values 1..4, a different one for every element) makes a
write that lands at the wrong offset or a tile taken from another step visible; with SGD lr=1, momentum 0 the master after step t must
equal  -pattern * sum_{s<=t} (s%5) * mean(k+1)  everywhere, and every replica must hold its bf16 rounding.  The replica is checked on the
DEVICE every step without a host synchronisation (a mismatch counter is accumulated by an asynchronous compare), so the check does not
serialise the protocol it is testing.  argv: topology steps [push|pull] [size: small|resnet18] [K]  (K < P: staggered workers, ps only)."""
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
big = len(sys.argv) > 4 and sys.argv[4] == "resnet18"
K = int(sys.argv[5]) if len(sys.argv) > 5 else world
T = 128 * 87297 if big else 128 * 257          # 11.17 M elements = the ResNet-18 arena
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
pattern = (torch.arange(T, device=dev) % 4 + 1).float()
expect = 0.0
bad = 0
nbad = torch.zeros((), dtype=torch.int64, device=dev)
if K == world:
    for s in range(1, steps + 1):
        if master is not None:
            tr.ps_publish(s, lowp)
        tr.worker_fetch(lowp)
        # the replica every worker computes with must be the result of step s-1 (bf16-rounded); compared on the device, no host sync
        nbad += (lowp.flat != (pattern * expect).to(torch.bfloat16)).sum()
        grads.flat.copy_(pattern * float((tr.worker_slot(rank) + 1) * (s % 5)))
        tr.worker_push(s, grads)
        if master is not None:
            tr.ps_reduce_update(s, world, opt, master, lowp=lowp)
        expect -= (s % 5) * mean_k
    torch.cuda.synchronize()
    tr.check_error()
    bad += int(nbad.item())
    if master is not None:
        lo, hi = (0, T) if topology == "ps" else tr.shards[rank]
        if not bool((master.flat[lo:hi] == pattern[lo:hi] * expect).all()):
            bad += 100
else:
    # K of P with staggered arrivals: worker r delays its push by r device-milliseconds, so different workers win in different steps.  Every
    # element of one update must come from ONE set of K workers: (w_prev - w_new) / (pattern * (s%5)) must be the same number everywhere, and
    # that number must be the mean of K distinct values out of {1..P}.  A worker whose push came too late finds the server one or more steps ahead,
    # adopts its step (the payload is built from the step worker_fetch returns) and the job ends with the server's shutdown publication.
    import itertools
    legal = sorted({sum(c) / K for c in itertools.combinations(range(1, world + 1), K)})
    prev = torch.zeros(T, device=dev)
    tr.follow_server = True            # K < P: a worker that is left behind adopts the server's step; worker_fetch returns the step it computes
    it = 0
    while True:
        it += 1
        if master is not None:
            if it > steps:
                tr.ps_shutdown()       # releases everybody who still waits for a publication
                break
            tr.ps_publish(it, lowp)
        d = tr.worker_fetch(lowp)
        if d < 0 or d > steps:
            break
        torch.cuda._sleep(int(1.5e6 * ((rank + it) % world)))
        grads.flat.copy_(pattern * float((tr.worker_slot(rank) + 1) * (1 + d % 3)))
        tr.worker_push(d, grads)
        if master is not None:
            assert d == it, (d, it)
            tr.ps_reduce_update(it, K, opt, master, lowp=lowp)
            torch.cuda.synchronize()
            q = (prev - master.flat) / (pattern * float(1 + it % 3))
            qlo, qhi = float(q.min()), float(q.max())
            # 1/K is not a binary fraction for K = 3: allow rounding noise (1e-4), nothing near the 1/K = 0.33 that separates two worker sets
            if qhi - qlo > 1e-4 or min(abs(qlo - v) for v in legal) > 1e-4:
                bad += 1
            prev = master.flat.clone()
    torch.cuda.synchronize()
    tr.check_error()
t = torch.tensor([bad])
dist.all_reduce(t)
if rank == 0:
    print("LITMUS", topology, "pull" if pull else "push", "steps", steps, "mismatches", int(t[0]))
dist.barrier()
sys.exit(1 if int(t[0]) else 0)
