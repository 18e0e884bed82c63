"""Same-process transport for the legacy runners (reference engine/transport/local.py:11-22).

A node registers a delivery callback; ``send`` looks the callback up and invokes it synchronously on
the caller's thread.  It also keeps per-node delivery counters, handy in tests and demos.
"""
from __future__ import annotations

import collections
import threading
from typing import Any, Callable, Dict, Iterable


class LocalTransport:
    """Same-process :class:`~byzpy_b200.engine.transport.base.Transport`: ``send`` calls the registered handler on the
    caller's thread.  ``delivered`` counts deliveries per node; sending to an unknown node raises ``KeyError``.

    Examples
    --------
    >>> from byzpy_b200.engine.transport.local import LocalTransport
    >>> t, seen = LocalTransport(), []
    >>> t.register("n0", seen.append)
    >>> t.send("n0", {"x": 1}); seen, t.delivered["n0"]
    ([{'x': 1}], 1)
    """

    def __init__(self) -> None:
        self._lock = threading.Lock()
        self._deliver: Dict[str, Callable[[Any], None]] = {}
        self.delivered: "collections.Counter[str]" = collections.Counter()

    def register(self, node_id: str, handler: Callable[[Any], None]) -> None:
        if not callable(handler):
            raise TypeError("handler must be callable")
        with self._lock:
            self._deliver[node_id] = handler

    def unregister(self, node_id: str) -> None:
        with self._lock:
            self._deliver.pop(node_id, None)

    def known_nodes(self) -> Iterable[str]:
        with self._lock:
            return tuple(self._deliver)

    def send(self, to_id: str, payload: Any) -> None:
        with self._lock:
            callback = self._deliver.get(to_id)
        if callback is None:
            raise KeyError(f"Unknown node_id {to_id}")
        callback(payload)
        self.delivered[to_id] += 1


__all__ = ["LocalTransport"]
