"""Payload codecs of the CPU/gloo tier.

``compress``/``decompress_into`` are the lossless path that plays the role of the reference's Blosc/snappy
``g_compress``/``w_compress`` (``/root/reference/src/compression.py:18-46``): a byte-shuffle (split each
fp32 into its 4 byte planes — exponent bytes of neighbouring gradients are highly repetitive) followed by an
LZ77-class coder.  The coder is the in-tree C++ extension (``ops/csrc/lz_codec.cpp``) when it is built, with a
pure-python/zlib fallback so the plumbing tier never needs a compiler.  On NVLink the replacement for Blosc is
the in-kernel block-scaled fp8 cast (``ops/csrc/comm_kernels.cu``), not this file.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np
import torch

_MAGIC = 0x50533230  # "PS20"


def _native():
    try:
        from ..ops import host_ext
        return host_ext.get()
    except Exception:
        return None


def compress(t: torch.Tensor) -> torch.Tensor:
    """Lossless: tensor → uint8 blob (header: magic, codec id, itemsize, raw bytes)."""
    isz = t.element_size()
    raw = t.detach().contiguous().reshape(-1).view(torch.uint8) if t.numel() else torch.empty(0, dtype=torch.uint8)
    nat = _native()
    if nat is not None:
        body = nat.shuffle_lz_compress(raw, isz)
        cid = 1
    else:
        a = raw.numpy().reshape(-1, isz).T.copy()          # byte planes
        body = torch.from_numpy(np.frombuffer(zlib.compress(a.tobytes(), 1), dtype=np.uint8).copy())
        cid = 2
    hdr = torch.from_numpy(np.frombuffer(struct.pack("<IIQQ", _MAGIC, cid, isz, raw.numel()), dtype=np.uint8).copy())
    return torch.cat([hdr, body])


def decompress_into(blob: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    magic, cid, isz, nraw = struct.unpack("<IIQQ", bytes(blob[:24].numpy().tobytes()))
    assert magic == _MAGIC, "corrupt payload"
    dst = out.reshape(-1).view(torch.uint8) if out.numel() else torch.empty(0, dtype=torch.uint8)
    assert dst.numel() == nraw, (dst.numel(), nraw)
    if nraw == 0:
        return out
    body = blob[24:]
    if cid == 1:
        nat = _native()
        assert nat is not None, "payload was produced by the native codec, which is not built here"
        nat.shuffle_lz_decompress(body, dst, isz)
    else:
        planes = np.frombuffer(zlib.decompress(body.numpy().tobytes()), dtype=np.uint8).reshape(isz, -1)
        dst.copy_(torch.from_numpy(np.ascontiguousarray(planes.T).reshape(-1)))
    return out


def encode_grad(flat: torch.Tensor, use_compress: bool) -> torch.Tensor:
    g = flat.detach().to(torch.float32).contiguous()
    if use_compress:
        return compress(g)
    return g.view(torch.uint8).reshape(-1).clone()


def decode_grad(blob: torch.Tensor, out_f32: torch.Tensor, use_compress: bool) -> None:
    if use_compress:
        decompress_into(blob, out_f32)
    else:
        out_f32.view(torch.uint8).reshape(-1).copy_(blob)


# API-compatible names of the reference module ------------------------------------------------------------
def g_compress(grad):
    return bytes(compress(torch.as_tensor(np.ascontiguousarray(grad))).numpy().tobytes()) + _shape_trailer(grad)


def _shape_trailer(a) -> bytes:
    a = np.asarray(a)
    dt = a.dtype.str.encode()
    return struct.pack("<B", len(dt)) + dt + struct.pack("<B", a.ndim) + struct.pack("<%dq" % a.ndim, *a.shape) + \
        struct.pack("<I", 1 + len(dt) + 1 + 8 * a.ndim)


def g_decompress(msg: bytes):
    assert isinstance(msg, (bytes, bytearray))
    tl = struct.unpack("<I", msg[-4:])[0]
    tr = msg[-4 - tl:-4]
    ld = tr[0]
    dt = np.dtype(tr[1:1 + ld].decode())
    nd = tr[1 + ld]
    shape = struct.unpack("<%dq" % nd, tr[2 + ld:2 + ld + 8 * nd])
    out = torch.empty(int(np.prod(shape)) if nd else 1, dtype=torch.from_numpy(np.zeros(1, dt)).dtype)
    blob = torch.from_numpy(np.frombuffer(msg[:-4 - tl], dtype=np.uint8).copy())
    decompress_into(blob, out)
    return out.numpy().reshape(shape)


w_compress, w_decompress = g_compress, g_decompress
