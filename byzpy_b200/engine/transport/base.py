"""``Transport`` protocol: ``register(node_id, handler)`` / ``send(to_id, payload)``
(reference engine/transport/base.py:9-16)."""
from __future__ import annotations

from typing import Any, Callable, Protocol


class Transport(Protocol):
    """Structural type of the legacy runners' message transports: ``register(node_id, handler)`` announces where to
    deliver a node's messages, ``send(to_id, payload)`` delivers one.  :class:`~byzpy_b200.engine.transport.local.
    LocalTransport` and :class:`~byzpy_b200.engine.transport.tcp.TcpTransport` implement it.
    """

    def register(self, node_id: str, handler: Callable[[Any], None]) -> None: ...

    def send(self, to_id: str, payload: Any) -> None: ...


__all__ = ["Transport"]
