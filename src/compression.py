"""API-compatible facade (reference ``src/compression.py``): lossless codec of the CPU tier.  On NVLink the
gradient "compression" is the in-kernel block-scaled fp8 cast (``ps_pytorch_b200/ops/csrc/comm_kernels.cu``)."""
import _path  # noqa: F401
from ps_pytorch_b200.utils.codec import g_compress, g_decompress, w_compress, w_decompress  # noqa: F401
