"""API-compatible facade (reference ``src/sync_replicas_master_nn.py``)."""
import _path  # noqa: F401
from ps_pytorch_b200.parallel.ps import SyncReplicasMaster_NN, GradientAccumulator, STEP_START_  # noqa: F401
from ps_pytorch_b200.utils.metrics import accuracy  # noqa: F401
