"""Actor endpoints and channel handles (reference engine/actor/channels.py:13-65)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional

from .ipc import unwrap_payload


@dataclass(frozen=True)
class Endpoint:
    """Globally addressable actor location."""

    scheme: str    # "thread" | "process" | "gpu" | "tcp" | "ucx"
    address: str   # "" for in-host schemes, "host:port" for tcp / ucx
    actor_id: str  # unique within (scheme, address)


class ChannelRef:
    """Handle on the mailbox ``name`` of a local actor: ``send(to, payload)`` / ``recv(timeout)``."""

    __slots__ = ("_backend", "_local", "_name")

    def __init__(self, backend, local_ep: Endpoint, name: str):
        self._backend = backend
        self._local = local_ep
        self._name = name

    @property
    def endpoint(self) -> Endpoint:
        return self._local

    @property
    def name(self) -> str:
        return self._name

    async def send(self, to: Endpoint, payload: Any) -> None:
        await self._backend.chan_put(from_ep=self._local, to_ep=to, name=self._name, payload=payload)

    async def recv(self, *, timeout: Optional[float] = None) -> Any:
        raw = await self._backend.chan_get(ep=self._local, name=self._name, timeout=timeout)
        return unwrap_payload(raw)


async def open_channel(backend, name: str) -> ChannelRef:
    return ChannelRef(backend, await backend.chan_open(name), name)


__all__ = ["Endpoint", "ChannelRef", "open_channel"]
