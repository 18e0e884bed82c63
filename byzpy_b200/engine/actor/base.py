"""What an actor backend must provide, and the ``ActorRef`` handle user code holds.

Contract of reference engine/actor/base.py:8-60.  A backend hosts ONE object (built by
``construct``) somewhere -- a worker thread, a spawned process, a CUDA-stream worker, a remote TCP
server -- executes method calls on it, and owns named mailboxes ("channels") other actors can post
to.  ``ActorRef`` turns attribute access into awaitable remote calls.
"""
from __future__ import annotations

import functools
from typing import Any, Optional, Protocol, runtime_checkable

from .channels import ChannelRef, Endpoint


@runtime_checkable
class ActorBackend(Protocol):
    """Structural type: the thread / process / gpu / remote backends do not inherit from it."""

    # lifecycle
    async def start(self) -> None: ...
    async def close(self) -> None: ...

    # the hosted object
    async def construct(self, cls_or_factory: Any, *, args: tuple, kwargs: dict) -> None: ...
    async def call(self, method: str, *args, **kwargs) -> Any: ...

    # addressing + mailboxes
    async def get_endpoint(self) -> Endpoint: ...
    async def chan_open(self, name: str) -> Endpoint: ...
    async def chan_put(self, *, from_ep: Endpoint, to_ep: Endpoint, name: str, payload: Any) -> None: ...
    async def chan_get(self, *, ep: Endpoint, name: str, timeout: Optional[float]) -> Any: ...


async def _invoke(backend: ActorBackend, method: str, *args, **kwargs) -> Any:
    return await backend.call(method, *args, **kwargs)


class ActorRef:
    """``await ref.some_method(x)`` runs ``some_method(x)`` on the hosted object.

    Also an async context manager (``async with ref:`` starts / closes the backend).  Dunder
    lookups are never forwarded, so pickling / copying a ref does not trigger remote calls.
    """

    def __init__(self, backend: ActorBackend):
        self._backend = backend

    def __getattr__(self, name: str):
        if name[:2] == "__" == name[-2:]:
            raise AttributeError(name)
        call = functools.partial(_invoke, self._backend, name)
        functools.update_wrapper(call, _invoke)
        call.__name__ = name
        return call

    async def __aenter__(self) -> "ActorRef":
        await self._backend.start()
        return self

    async def __aexit__(self, exc_type, exc, tb) -> bool:
        await self._backend.close()
        return False

    async def endpoint(self) -> Endpoint:
        return await self._backend.get_endpoint()

    async def open_channel(self, name: str) -> ChannelRef:
        ep = await self._backend.chan_open(name)
        return ChannelRef(self._backend, ep, name)


__all__ = ["ActorBackend", "ActorRef"]
