"""Compatibility surface of the reference's UCX transport module
(reference engine/actor/transports/ucx.py).  There is no UCX dependency here: on one NVSwitch box
the GPU-direct tensor path of the ``ucx://`` scheme is CUDA IPC (``transports/cuda_ipc.py``)."""
from __future__ import annotations

from typing import Any, Tuple

from . import cuda_ipc


def _ucx_mod():
    """The installed UCX Python binding (``ucxx``, else ``ucp``), or None -- the probe the reference's
    callers use (reference engine/actor/transports/ucx.py:36-56).  Nothing here needs it: the ``ucx://``
    scheme moves device tensors as CUDA-IPC handles over the TCP control plane."""
    import importlib

    for name in ("ucxx", "ucp"):
        try:
            return importlib.import_module(name)
        except Exception:  # noqa: BLE001  (absent, or present but unusable on this box)
            continue
    return None


def have_ucx() -> bool:
    """True when the GPU-direct payload path is usable (a CUDA device is present)."""
    return cuda_ipc.available()


def pack_payload(obj: Any) -> Tuple[str, bytes]:
    blob = cuda_ipc.dumps(obj, same_host=True)
    return ("cuda" if blob[:1] == b"I" else "pickle"), blob


def unpack_payload(blob: bytes) -> Any:
    return cuda_ipc.loads(blob)


__all__ = ["have_ucx", "pack_payload", "unpack_payload"]
