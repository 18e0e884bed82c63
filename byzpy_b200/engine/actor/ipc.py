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

from ..storage.shared_store import SharedTensorHandle, cleanup_tensor, open_tensor, register_tensor

_SHM_MARK = "__BYZ_SHARED_TENSOR__"


def wrap_payload(obj: Any) -> Any:
    if isinstance(obj, torch.Tensor):
        return (_SHM_MARK, register_tensor(obj.detach().cpu().numpy()))
    if isinstance(obj, np.ndarray):
        return (_SHM_MARK, register_tensor(obj))
    if isinstance(obj, tuple) and len(obj) == 2 and obj[0] == _SHM_MARK:
        return obj
    if isinstance(obj, tuple) and hasattr(obj, "_fields"):      # namedtuple (e.g. an Endpoint)
        return type(obj)(*(wrap_payload(x) for x in obj))
    if isinstance(obj, (list, tuple)):
        return type(obj)(wrap_payload(x) for x in obj)
    if isinstance(obj, dict):
        return {k: wrap_payload(v) for k, v in obj.items()}
    return obj


def _take(handle: SharedTensorHandle) -> torch.Tensor:
    with open_tensor(handle) as arr:
        data = np.array(arr, copy=True)
    cleanup_tensor(handle)
    return torch.from_numpy(data)


def unwrap_payload(obj: Any) -> Any:
    if isinstance(obj, tuple) and len(obj) == 2 and obj[0] == _SHM_MARK:
        return _take(obj[1])
    if isinstance(obj, list):
        return [unwrap_payload(x) for x in obj]
    if isinstance(obj, tuple) and hasattr(obj, "_fields"):
        return type(obj)(*(unwrap_payload(x) for x in obj))
    if isinstance(obj, tuple):
        return tuple(unwrap_payload(x) for x in obj)
    if isinstance(obj, dict):
        return {k: unwrap_payload(v) for k, v in obj.items()}
    return obj


__all__ = ["wrap_payload", "unwrap_payload"]
