"""BASELINE.json config 1: LeNet / MNIST-shaped synthetic data, 1 PS + 1 worker on CPU over gloo (plumbing, no GPU).
    python -m ps_pytorch_b200.launch -n 2 scripts/bench_cpu_plumbing.py [--steps 50] [--batch 64] [--compress 0|1]
Rank 0 prints one JSON line: images/s through the full PS path (weights bcast -> fwd/bwd -> gradient push -> aggregate -> SGD)
next to the same model stepped by a plain single-process torch loop on the same machine."""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200.engine import PSJob
from ps_pytorch_b200.models import build_model

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=50); ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--batch", type=int, default=64); ap.add_argument("--compress", type=int, default=0)
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.set_num_threads(int(os.environ.get("OMP_NUM_THREADS", "4")))
dist.init_process_group("gloo", rank=rank, world_size=world)
job = PSJob(network="LeNet", dataset="MNIST", batch_size=a.batch, device="cpu", dtype="fp32", lr=0.05, momentum=0.9, transport="gloo",
            colocate=False, grad_dtype="fp32", compress=bool(a.compress), cuda_graph=False)
x, y = torch.randn(a.batch, 1, 28, 28), torch.randint(0, 10, (a.batch,))
for _ in range(a.warmup):
    job.step(x, y)
dist.barrier(); t0 = time.time()
for _ in range(a.steps):
    job.step(x, y)
dist.barrier(); dt = time.time() - t0
job.close()
if rank == 0:
    m = build_model("LeNet", 10); opt = torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9); crit = torch.nn.CrossEntropyLoss()
    for i in range(a.warmup + a.steps):
        if i == a.warmup: t1 = time.time()
        opt.zero_grad(); crit(m(x), y).backward(); opt.step()
    ds = time.time() - t1
    print(json.dumps({"config": "LeNet/MNIST-shaped synthetic, 1 PS + 1 worker, CPU/gloo", "batch": a.batch, "steps": a.steps, "compress": bool(a.compress),
                      "ps_images_per_s": a.batch * a.steps / dt, "ps_ms_per_step": dt / a.steps * 1e3,
                      "single_process_images_per_s": a.batch * a.steps / ds, "single_process_ms_per_step": ds / a.steps * 1e3,
                      "threads": torch.get_num_threads()}))
dist.barrier()
