"""In-process transport: ``send`` calls the registered handler directly
(reference engine/transport/local.py:11-22)."""
from __future__ import annotations

from typing import Any, Callable, Dict


class LocalTransport:
    def __init__(self) -> None:
        self._handlers: Dict[str, Callable[[Any], None]] = {}

    def register(self, node_id: str, handler: Callable[[Any], None]) -> None:
        self._handlers[node_id] = handler

    def send(self, to_id: str, payload: Any) -> None:
        try:
            handler = self._handlers[to_id]
        except KeyError:
            raise KeyError(f"Unknown node_id {to_id}") from None
        handler(payload)


__all__ = ["LocalTransport"]
