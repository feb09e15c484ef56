"""API-compatible facade (reference ``src/distributed_worker.py``)."""
import _path  # noqa: F401
from ps_pytorch_b200.parallel.worker import DistributedWorker, ModelBuffer, STEP_START_  # noqa: F401
from ps_pytorch_b200.utils.metrics import accuracy  # noqa: F401
