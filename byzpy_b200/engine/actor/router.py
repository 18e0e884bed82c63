"""Process-local registry ``(scheme, actor_id) -> backend`` so that any backend can deliver a
channel message to any in-process peer without importing its module
(reference engine/actor/router.py:17-55)."""
from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple


@dataclass(frozen=True)
class BackendRecord:
    """One row of the :class:`ChannelRouter` table: which backend object hosts actor ``actor_id`` of ``scheme``.
    """

    scheme: str
    actor_id: str
    backend: Any


class ChannelRouter:
    """Process-wide registry of the actor backends living in this process.

    Backends register themselves under ``(scheme, actor_id)``; when one of them must deliver a channel message to an
    endpoint, ``resolve`` tells it whether the addressee is in this process (then the payload is put on its asyncio
    queue directly) or has to be reached through a transport.  Thread safe.  The module-level ``channel_router`` is the
    instance everybody uses.

    Examples
    --------
    >>> from byzpy_b200.engine.actor.router import ChannelRouter
    >>> r = ChannelRouter()
    >>> r.register("thread", "a1", backend := object())
    >>> r.resolve("thread", "a1") is backend, r.resolve("thread", "nobody")
    (True, None)
    """

    def __init__(self) -> None:
        self._lock = threading.Lock()
        self._table: Dict[Tuple[str, str], BackendRecord] = {}

    def register(self, scheme: str, actor_id: str, backend: Any) -> None:
        with self._lock:
            self._table[(scheme, actor_id)] = BackendRecord(scheme, actor_id, backend)

    def unregister(self, scheme: str, actor_id: str) -> None:
        with self._lock:
            self._table.pop((scheme, actor_id), None)

    def resolve(self, scheme: str, actor_id: str) -> Optional[Any]:
        rec = self._table.get((scheme, actor_id))
        return None if rec is None else rec.backend


channel_router = ChannelRouter()

__all__ = ["ChannelRouter", "BackendRecord", "channel_router"]
