"""fp8 gradient wire vs bf16 vs fp32: loss curves of the same job (ResNet-18, synthetic class-separable CIFAR-shaped data,
batch 256, SGD 0.05/0.9) through the full PS path (push kernel -> reduce_update kernel) on one GPU.  Evidence for the
lossy block-scaled fp8 cast that replaces the reference's lossless Blosc codec (SURVEY.md §7.5 risk 7)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ps_pytorch_b200.engine import PSJob  # noqa: E402
from ps_pytorch_b200.utils.data import synthetic_dataset, BatchLoader  # noqa: E402

steps = int(os.environ.get("STEPS", 80))
out = {}
for wire in ("bf16", "fp8"):
    torch.manual_seed(0)
    job = PSJob(network="ResNet18", dataset="Cifar10", batch_size=256, device="cuda:0", dtype="bf16", lr=0.05, momentum=0.9,
                grad_dtype=wire, cuda_graph=False, backend="sm100", seed=3)
    ds = synthetic_dataset("Cifar10", 8192, seed=0)
    loader = BatchLoader(ds, 256, shuffle=True, augment=True, seed=5, pin=True, channels_last=True, prefetch=0)
    losses, it = [], iter(loader)
    for s in range(steps):
        try:
            x, y = next(it)
        except StopIteration:
            it = iter(loader); x, y = next(it)
        losses.append(float(job.step(x, y)))
    job.close()
    out[wire] = losses
    print(wire, "loss[0,10,20,40,-1] =", [round(losses[i], 4) for i in (0, 10, 20, 40, -1)], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/convergence.json", "w"))
