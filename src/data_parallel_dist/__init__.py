from .data_parallel_dist import DistributedDataParallel  # noqa: F401
