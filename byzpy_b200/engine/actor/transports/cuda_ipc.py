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
import os
import pickle
import uuid
import weakref
from multiprocessing.reduction import ForkingPickler
from typing import Any

import torch


def available() -> bool:
    """CUDA-IPC payloads need a CUDA device in this process."""
    return torch.cuda.is_available()


# CUDA IPC cannot open a handle inside the process that exported it.  A payload that comes back to its
# sender (a tensor parked in a remote mailbox and fetched by the same process, a relay that echoes)
# therefore resolves against this registry of live exported tensors instead of the handle.
_EXPORTED: "weakref.WeakValueDictionary[str, torch.Tensor]" = weakref.WeakValueDictionary()


def _rebuild(pid: int, key: str, fn, args):
    if pid == os.getpid():
        t = _EXPORTED.get(key)
        if t is not None:
            return t
    return fn(*args)


class _IpcPickler(ForkingPickler):
    def reducer_override(self, obj):
        if isinstance(obj, torch.Tensor) and obj.is_cuda:
            from torch.multiprocessing.reductions import reduce_tensor

            fn, args = reduce_tensor(obj)
            key = uuid.uuid4().hex
            _EXPORTED[key] = obj
            return _rebuild, (os.getpid(), key, fn, args)
        return NotImplemented


def dumps(obj: Any, *, same_host: bool = True) -> bytes:
    """Serialise ``obj``; CUDA tensors become IPC handles when ``same_host``."""
    if same_host and available():
        import torch.multiprocessing  # noqa: F401  (registers the CUDA reductions)

        buf = io.BytesIO()
        _IpcPickler(buf, pickle.HIGHEST_PROTOCOL).dump(obj)
        return b"I" + buf.getvalue()
    return b"V" + pickle.dumps(_to_host(obj), protocol=pickle.HIGHEST_PROTOCOL)


_ctx_ready = False


def _ensure_context() -> None:
    """Opening an IPC handle needs a live CUDA context on the calling thread; a receiver process that
    has not touched its GPU yet (an actor server that only relays) gets one here."""
    global _ctx_ready
    if not _ctx_ready and available():
        torch.empty(0, device="cuda")
        _ctx_ready = True
    if available():
        torch.cuda.set_device(torch.cuda.current_device())      # bind the primary context to THIS thread


def loads(data: bytes) -> Any:
    """Rebuild a payload written by :func:`dumps`: ``b"I"`` frames map the sender's device memory (CUDA-IPC handles),
    ``b"V"`` frames carry host copies.
    """
    tag, body = data[:1], data[1:]
    if tag == b"I":
        _ensure_context()
        return pickle.loads(body)
    if tag == b"V":
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
