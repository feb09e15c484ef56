"""API-compatible facade (reference ``src/util.py``): ``build_model`` + ``prepare_data``."""
import _path  # noqa: F401
from ps_pytorch_b200.models import build_model  # noqa: F401
from ps_pytorch_b200.utils.data import prepare_data  # noqa: F401
