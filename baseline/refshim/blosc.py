"""Stand-in for python-blosc's ``pack_array`` / ``unpack_array`` (the only two calls the reference makes,
/root/reference/src/compression.py:18-46).  python-blosc is not installed and there is no network.  The container is a
small header + the raw array bytes (no shuffle/snappy pass at all), which is never slower than the real codec on float32
gradients — i.e. it favours the reference arm."""
import pickle
import struct

import numpy as np

_MAGIC = b"RBL1"


def pack_array(array, clevel=9, shuffle=1, cname="blosclz"):
    a = np.ascontiguousarray(array)
    meta = pickle.dumps((a.dtype.str, a.shape), protocol=2)
    return _MAGIC + struct.pack("<I", len(meta)) + meta + a.tobytes()


def unpack_array(packed):
    assert packed[:4] == _MAGIC, "not a refshim blosc container"
    (n,) = struct.unpack("<I", packed[4:8])
    dtype, shape = pickle.loads(packed[8:8 + n])
    return np.frombuffer(packed, dtype=np.dtype(dtype), offset=8 + n).reshape(shape).copy()
