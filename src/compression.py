"""API-compatible facade (reference ``src/compression.py``): lossless codec of the CPU tier.  On NVLink the
gradient "compression" is the in-kernel block-scaled fp8 cast (``ps_pytorch_b200/ops/csrc/comm_kernels.cu``)."""
import _path  # noqa: F401
from ps_pytorch_b200.utils.codec import g_compress, g_decompress, w_compress, w_decompress  # noqa: F401

_SENTINEL = bytes([0x29]) * 32


def _trim_msg(msg):
    """Cut a fixed-size receive buffer at the 32×0x29 terminator (reference ``src/compression.py:8-16``; unused there as well, since frames
    carry their own length — kept so legacy callers resolve)."""
    msg = bytes(msg)
    i = msg.find(_SENTINEL)
    return msg if i < 0 else msg[:i]
