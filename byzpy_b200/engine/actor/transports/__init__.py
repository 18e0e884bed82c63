"""Transports behind the actor channels: ``tcp`` (control plane, any host) and ``cuda_ipc``
(GPU-direct payloads on one NVSwitch box; the B200-native stand-in for the reference's UCX
transport)."""
from . import cuda_ipc, tcp, ucx

__all__ = ["tcp", "ucx", "cuda_ipc"]
