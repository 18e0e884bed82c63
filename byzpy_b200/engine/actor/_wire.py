"""Length-prefixed object framing for the TCP control plane: 4-byte big-endian length +
cloudpickle body (frame layout of reference engine/actor/_wire.py:5-18; cloudpickle instead of
pickle so classes/closures defined in ``__main__`` travel by value)."""
from __future__ import annotations

import asyncio
import pickle
import struct
from typing import Any

import cloudpickle

_HDR = struct.Struct(">I")
MAX_FRAME = (1 << 32) - 1


def encode(obj: Any) -> bytes:
    """``obj`` as one frame: 4-byte big-endian length + cloudpickle body."""
    body = cloudpickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    if len(body) > MAX_FRAME:
        raise ValueError("frame too large")
    return _HDR.pack(len(body)) + body


async def send_obj(w: asyncio.StreamWriter, obj: Any) -> None:
    """Write ``obj`` as one frame to the stream writer ``w`` and drain."""
    w.write(encode(obj))
    await w.drain()


async def recv_obj(r: asyncio.StreamReader) -> Any:
    """Read one frame from the stream reader ``r`` and unpickle it."""
    (n,) = _HDR.unpack(await r.readexactly(_HDR.size))
    return cloudpickle.loads(await r.readexactly(n))


__all__ = ["send_obj", "recv_obj", "encode"]
