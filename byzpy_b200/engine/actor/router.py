"""Process-local registry ``(scheme, actor_id) -> backend`` so that any backend can deliver a
channel message to any in-process peer without importing its module
(reference engine/actor/router.py:17-55)."""
from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple


@dataclass(frozen=True)
class BackendRecord:
    scheme: str
    actor_id: str
    backend: Any


class ChannelRouter:
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
