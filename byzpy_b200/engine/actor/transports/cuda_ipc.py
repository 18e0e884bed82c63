"""GPU-direct payload codec for the ``ucx://`` scheme on one NVSwitch box.

The reference ships CUDA tensors between actor servers through UCX tagged sends of raw device
buffers (reference engine/actor/transports/ucx.py:186-277; requires ucxx + cupy).  On a single
B200 node every GPU reaches every other at full NVLink bandwidth, so the native equivalent needs
no copy engine in the middle: the sender exports the tensor's allocation as a CUDA IPC handle
(``torch.multiprocessing`` reductions) and the receiver maps it -- zero-copy, device-to-device.
Payloads addressed to another host fall back to by-value serialisation.
"""
from __future__ import annotations

import io
import pickle
from multiprocessing.reduction import ForkingPickler
from typing import Any

import torch


def available() -> bool:
    return torch.cuda.is_available()


def dumps(obj: Any, *, same_host: bool = True) -> bytes:
    """Serialise ``obj``; CUDA tensors become IPC handles when ``same_host``."""
    if same_host and available():
        import torch.multiprocessing  # noqa: F401  (registers the CUDA reductions)

        buf = io.BytesIO()
        ForkingPickler(buf, pickle.HIGHEST_PROTOCOL).dump(obj)
        return b"I" + buf.getvalue()
    return b"V" + pickle.dumps(_to_host(obj), protocol=pickle.HIGHEST_PROTOCOL)


def loads(data: bytes) -> Any:
    tag, body = data[:1], data[1:]
    if tag in (b"I", b"V"):
        return pickle.loads(body)
    raise ValueError("unknown cuda_ipc payload tag")


def _to_host(obj: Any) -> Any:
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu()
    if isinstance(obj, tuple) and hasattr(obj, "_fields"):      # namedtuple (e.g. channels.Endpoint)
        return type(obj)(*(_to_host(x) for x in obj))
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_host(x) for x in obj)
    if isinstance(obj, dict):
        return {k: _to_host(v) for k, v in obj.items()}
    return obj


__all__ = ["dumps", "loads", "available"]
