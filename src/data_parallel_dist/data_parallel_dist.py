"""All-reduce data parallelism under the reference's module path (``/root/reference/src/data_parallel_dist/data_parallel_dist.py:29-267``,
a vendored torch-0.2 DistributedDataParallel that no entry point used).

Here it is a live mode of the framework: ``ps_pytorch_b200.parallel.ddp.DistributedDataParallel`` broadcasts rank 0's parameters at
construction (reference ``:46,136``) and averages gradients with ONE launch of the PS reduce kernel in all-reduce mode over the symmetric
NVLink arena (``reduce_update_kernel`` kind = 2) — or ``torch.distributed.all_reduce`` on the CPU/gloo tier — instead of 1 MB NCCL buckets
(reference ``:74,252,261``).  Select it from the CLI with ``--parallelism ddp``.
"""
import _path  # noqa: F401  (puts the repository root on sys.path when run from src/)
from ps_pytorch_b200.parallel.ddp import DistributedDataParallel  # noqa: F401

__all__ = ["DistributedDataParallel"]
