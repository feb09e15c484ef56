"""API-compatible facade (reference ``src/nn_ops.py``)."""
import _path  # noqa: F401
from ps_pytorch_b200.parallel.trainer import NN_Trainer  # noqa: F401
from ps_pytorch_b200.utils.metrics import accuracy  # noqa: F401
