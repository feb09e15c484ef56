"""Multi-process helper (run under ps_pytorch_b200.launch): a PS job with deterministic per-worker batches.
Rank 0 writes the final master weights to --out so the test can compare with a single-process reference."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200.engine import PSJob  # noqa: E402


def batch_for(worker_slot, step, bs, shape, nc):
    g = torch.Generator().manual_seed(1000 * worker_slot + step)
    return torch.randn(bs, *shape, generator=g), torch.randint(0, nc, (bs,), generator=g)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--network", default="LeNet")
    ap.add_argument("--transport", default="gloo")
    ap.add_argument("--topology", default="ps")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--grad-dtype", default="fp32")
    ap.add_argument("--colocate", type=int, default=0)
    ap.add_argument("--compress", type=int, default=0)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--bs", type=int, default=8)
    ap.add_argument("--overlap", type=int, default=0)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--backend", default="auto")
    ap.add_argument("--comm-type", default="Bcast")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if a.device == "cuda":
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        a.device = "cuda:%d" % int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = "MNIST" if a.network.startswith("LeNet") else "Cifar10"
    shape = (1, 28, 28) if ds == "MNIST" else (3, 32, 32)
    job = PSJob(network=a.network, dataset=ds, batch_size=a.bs, device=a.device, dtype=a.dtype, lr=0.05, momentum=0.9,
                transport=a.transport, topology=a.topology, colocate=bool(a.colocate), grad_dtype=a.grad_dtype,
                num_aggregate=(a.k or None), compress=bool(a.compress), overlap=bool(a.overlap), num_buckets=3,
                cuda_graph=bool(a.graph), backend=a.backend, seed=7, comm_type=a.comm_type)
    losses = []
    for s in range(1, a.steps + 1):
        x = y = None
        if job.is_worker:
            x, y = batch_for(job.comm.worker_slot(rank), s, a.bs, shape, 10)
        l = job.step(x, y)
        if l is not None and os.environ.get("PSB_DEBUG"): print("rank", rank, "step", s, float(l), flush=True)
        if l is not None:
            losses.append(float(l))
    if job.master is not None and hasattr(job.comm, "gather_master"):
        job.comm.gather_master(job.master.master)
    job.close()
    if job.master is not None and rank == 0:
        if a.device.startswith("cuda"):
            torch.cuda.synchronize()
        torch.save({"w": [v.detach().float().cpu().clone() for v in job.master.master.views], "losses": losses,
                    "workers": job.comm.num_workers, "stats": dict(job.comm.stats)}, a.out)
    dist.barrier()


if __name__ == "__main__":
    main()
