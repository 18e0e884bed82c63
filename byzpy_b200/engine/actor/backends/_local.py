"""Shared core of the in-host actor backends (thread / gpu / process parents).

Every in-host backend owns named mailboxes (``asyncio.Queue`` per channel name), registers
itself with the process-wide :data:`channel_router`, and can deliver to any peer scheme:
in-host peers through the router, ``tcp://`` / ``ucx://`` peers through the wire transports.
``chan_get`` with a timeout returns ``None`` when nothing arrived (reference
engine/actor/backends/thread.py:133-137).
"""
from __future__ import annotations

import asyncio
import uuid
from typing import Any, Dict, Optional

from ..channels import Endpoint
from ..router import channel_router
from ..transports import cuda_ipc, tcp

IN_HOST_SCHEMES = ("thread", "process", "gpu")


class LocalMailboxBackend:
    scheme = "thread"

    def __init__(self) -> None:
        self._actor_id = str(uuid.uuid4())
        self._queues: Dict[str, asyncio.Queue] = {}
        self._loop: Optional[asyncio.AbstractEventLoop] = None
        channel_router.register(self.scheme, self._actor_id, self)

    # ------------------------------------------------------------------- endpoints
    async def get_endpoint(self) -> Endpoint:
        return Endpoint(scheme=self.scheme, address="", actor_id=self._actor_id)

    async def chan_open(self, name: str) -> Endpoint:
        self._queues.setdefault(name, asyncio.Queue())
        return await self.get_endpoint()

    def _unregister(self) -> None:
        channel_router.unregister(self.scheme, self._actor_id)

    # -------------------------------------------------------------------- delivery
    async def _deliver_local(self, name: str, from_ep: Optional[Endpoint], payload: Any) -> None:
        await self._queues.setdefault(name, asyncio.Queue()).put((from_ep, payload))

    async def chan_put(self, *, from_ep: Endpoint, to_ep: Endpoint, name: str, payload: Any) -> None:
        if to_ep.scheme in IN_HOST_SCHEMES:
            if to_ep.scheme == self.scheme and to_ep.actor_id == self._actor_id:
                await self._deliver_local(name, from_ep, payload)
                return
            peer = channel_router.resolve(to_ep.scheme, to_ep.actor_id)
            if peer is None:
                raise RuntimeError(f"no local {to_ep.scheme} actor {to_ep.actor_id}")
            await peer._deliver_local(name, from_ep, payload)
            return
        if to_ep.scheme == "tcp":
            host, port = tcp.parse_address(to_ep.address)
            await tcp.chan_put(host, port, to_ep.actor_id, name, payload)
            return
        if to_ep.scheme == "ucx":
            host, port = tcp.parse_address(to_ep.address)
            blob = cuda_ipc.dumps(payload, same_host=host in ("127.0.0.1", "localhost"))
            await tcp.chan_put(host, port, to_ep.actor_id, name, {"__cuda_ipc__": blob})
            return
        raise RuntimeError(f"{type(self).__name__} cannot route to {to_ep.scheme!r}")

    async def chan_get(self, *, ep: Endpoint, name: str, timeout: Optional[float]) -> Any:
        if ep.scheme == self.scheme and ep.actor_id == self._actor_id:
            q = self._queues.setdefault(name, asyncio.Queue())
            if timeout is None:
                _, payload = await q.get()
                return payload
            try:
                _, payload = await asyncio.wait_for(q.get(), timeout=timeout)
                return payload
            except asyncio.TimeoutError:
                return None
        if ep.scheme in ("tcp", "ucx"):
            host, port = tcp.parse_address(ep.address)
            got = await tcp.chan_get(host, port, ep.actor_id, name, timeout)
            if isinstance(got, dict) and "__cuda_ipc__" in got:
                return cuda_ipc.loads(got["__cuda_ipc__"])
            return got
        raise RuntimeError("Endpoint mismatch")


__all__ = ["LocalMailboxBackend", "IN_HOST_SCHEMES"]
