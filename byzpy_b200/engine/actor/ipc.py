"""Payload wrapping for process boundaries on the host path.

``wrap_payload`` replaces every tensor / ndarray in a nested payload by a tagged
``SharedTensorHandle`` (one copy into POSIX shm); ``unwrap_payload`` copies it back out and unlinks
the segment (single consumer) -- contract of reference engine/actor/ipc.py:20-56.  CUDA tensors
are staged through the host here; the device-to-device path is CUDA IPC
(:mod:`byzpy_b200.engine.actor.transports.cuda_ipc`, :mod:`byzpy_b200.parallel.symmetric`).
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from ..storage.shared_store import (SharedTensorHandle, cleanup_tensor, open_tensor, register_rows,
                                    register_tensor)

_SHM_MARK = "__BYZ_SHARED_TENSOR__"
_SHM_BATCH_MARK = "__BYZ_SHARED_TENSOR_BATCH__"
_BATCH_MIN_BYTES = 1 << 16


def _batchable(seq) -> bool:
    """A list / tuple of >= 2 CPU tensors of one shape and dtype (the honest gradients handed to a Byzantine
    node, the neighbour vectors of a gossip round): they travel as ONE segment instead of one each."""
    if len(seq) < 2 or not all(isinstance(t, torch.Tensor) for t in seq):
        return False
    first = seq[0]
    if first.is_cuda or first.numel() == 0 or first.numel() * first.element_size() * len(seq) < _BATCH_MIN_BYTES:
        return False
    if first.dtype not in (torch.float32, torch.float64, torch.float16, torch.int64, torch.int32, torch.uint8):
        return False
    return all(t.shape == first.shape and t.dtype == first.dtype and not t.is_cuda for t in seq)


def wrap_payload(obj: Any) -> Any:
    """Replace every tensor / array in a nested payload by a tagged shared-memory handle (lists of equally shaped CPU
    tensors share one segment) so the payload can cross a process boundary without pickling the data.
    """
    if isinstance(obj, torch.Tensor):
        return (_SHM_MARK, register_tensor(obj.detach().cpu().numpy()))
    if isinstance(obj, np.ndarray):
        return (_SHM_MARK, register_tensor(obj))
    if isinstance(obj, tuple) and len(obj) == 2 and obj[0] == _SHM_MARK:
        return obj
    if isinstance(obj, tuple) and len(obj) == 4 and obj[0] == _SHM_BATCH_MARK:
        return obj
    if isinstance(obj, tuple) and hasattr(obj, "_fields"):      # namedtuple (e.g. an Endpoint)
        return type(obj)(*(wrap_payload(x) for x in obj))
    if isinstance(obj, (list, tuple)):
        if type(obj) in (list, tuple) and _batchable(obj):
            return (_SHM_BATCH_MARK, register_rows(obj), type(obj) is tuple, tuple(obj[0].shape))
        return type(obj)(wrap_payload(x) for x in obj)
    if isinstance(obj, dict):
        return {k: wrap_payload(v) for k, v in obj.items()}
    return obj


def _take_batch(handle: SharedTensorHandle, as_tuple: bool, shape) -> Any:
    stacked = _take(handle)                                  # (n, numel): one copy out, one unlink
    rows = [r.reshape(shape) for r in stacked.unbind(0)]
    return tuple(rows) if as_tuple else rows


def _take(handle: SharedTensorHandle) -> torch.Tensor:
    with open_tensor(handle) as arr:
        data = np.array(arr, copy=True)
    cleanup_tensor(handle)
    return torch.from_numpy(data)


def unwrap_payload(obj: Any) -> Any:
    """Inverse of :func:`wrap_payload`: copy the data out of shared memory and unlink the segments (single consumer)."""
    if isinstance(obj, tuple) and len(obj) == 2 and obj[0] == _SHM_MARK:
        return _take(obj[1])
    if isinstance(obj, tuple) and len(obj) == 4 and obj[0] == _SHM_BATCH_MARK:
        return _take_batch(obj[1], obj[2], obj[3])
    if isinstance(obj, list):
        return [unwrap_payload(x) for x in obj]
    if isinstance(obj, tuple) and hasattr(obj, "_fields"):
        return type(obj)(*(unwrap_payload(x) for x in obj))
    if isinstance(obj, tuple):
        return tuple(unwrap_payload(x) for x in obj)
    if isinstance(obj, dict):
        return {k: unwrap_payload(v) for k, v in obj.items()}
    return obj


def discard_payload(obj: Any) -> None:
    """Release the shared-memory segments of a wrapped payload that nobody is going to unwrap (the reply of a
    call whose caller was cancelled / timed out): segments are single-consumer and named, so an unread one would
    stay in ``/dev/shm`` until the machine reboots."""
    if isinstance(obj, tuple) and len(obj) == 2 and obj[0] == _SHM_MARK:
        cleanup_tensor(obj[1])
    elif isinstance(obj, tuple) and len(obj) == 4 and obj[0] == _SHM_BATCH_MARK:
        cleanup_tensor(obj[1])
    elif isinstance(obj, (list, tuple)):
        for x in obj:
            discard_payload(x)
    elif isinstance(obj, dict):
        for v in obj.values():
            discard_payload(v)


__all__ = ["wrap_payload", "unwrap_payload", "discard_payload"]
