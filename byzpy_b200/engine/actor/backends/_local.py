"""Shared core of the in-host actor backends (thread / gpu / process parents).

Every in-host backend owns named mailboxes (``asyncio.Queue`` per channel name), registers
itself with the process-wide :data:`channel_router`, and can deliver to any peer scheme:
in-host peers through the router, ``tcp://`` / ``ucx://`` peers through the wire transports.
``chan_get`` with a timeout returns ``None`` when nothing arrived (reference
engine/actor/backends/thread.py:133-137).
"""
from __future__ import annotations

import asyncio
import os
import threading
import time
import uuid
from typing import Any, Dict, Optional

import torch

from ..channels import Endpoint
from ..router import channel_router
from ..transports import cuda_ipc, tcp

IN_HOST_SCHEMES = ("thread", "process", "gpu")


class IntraOpGovernor:
    """Splits the host's intra-op (OpenMP / MKL) threads between the thread actors that are
    executing a method at the same moment.

    Ten node actors running forward/backward concurrently, each opening 8-thread parallel
    regions on an 8-core host, oversubscribe it ~10x (measured: 10 SmallCNN replicas, 177 ms for
    the gradient phase against 10 x 13.7 ms run back to back).  While ``k >= 2`` calls are in
    flight every one of them gets ``max(1, base // k)`` threads; the original count comes back when
    the last one returns, so whatever runs next on the coordinator (the aggregator) sees the full
    machine.  The reference does this by hand in one place only (Multi-Krum's pool subtasks,
    aggregators/geometric_wise/krum.py:461-475).  ``BYZPY_INTRAOP_GOVERNOR=0`` turns it off.
    """

    def __init__(self) -> None:
        self._lock = threading.Lock()
        self._active = 0
        self._base: Optional[int] = None
        self._tls = threading.local()       # OpenMP's thread-count setting is per calling thread
        self._default: Optional[int] = None  # last value handed to torch.set_num_threads by any thread
        self._peak, self._peak_at = 0, 0.0   # highest concurrency seen within the last PEAK_WINDOW seconds
        self.enabled = os.environ.get("BYZPY_INTRAOP_GOVERNOR", "1") not in ("0", "false", "False")

    PEAK_WINDOW = 1.0

    def _share(self) -> int:
        # A round's calls arrive one after another within microseconds: sizing by the in-flight count
        # alone would hand the first arrival the whole machine and the second half of it.  The recent
        # peak makes a steady stream of k-wide rounds give every call base // k from the first one on,
        # and decays back to "alone -> everything" a second after the concurrency stops.
        now = time.monotonic()
        if self._active >= self._peak or now - self._peak_at > self.PEAK_WINDOW:
            self._peak, self._peak_at = self._active, now
        k = max(self._active, self._peak)
        return self._base if k <= 1 else max(1, self._base // k)

    def __enter__(self) -> "IntraOpGovernor":
        if self.enabled:
            with self._lock:
                if self._active == 0:        # nobody in flight: (re)learn the process-wide setting
                    self._base = torch.get_num_threads() if self._base is None else self._base
                self._active += 1
                want = self._share()
            if getattr(self._tls, "threads", None) != want:
                torch.set_num_threads(want)
                self._tls.threads = self._default = want
        return self

    def __exit__(self, *exc) -> None:
        if self.enabled:
            with self._lock:
                self._active -= 1
                last = self._active == 0
            if last and self._default != self._base:
                # set_num_threads also moves the default that threads created later start from: whoever
                # leaves last puts it back, even if its own share already was the full count
                torch.set_num_threads(self._base)
                self._tls.threads = self._default = self._base


intra_op_governor = IntraOpGovernor()


class LocalMailboxBackend:
    """Mailbox and addressing half of every in-host backend (thread, process, gpu).

    Keeps one asyncio queue per channel name, registers the actor with the process-wide
    :class:`~byzpy_b200.engine.actor.router.ChannelRouter`, and implements ``chan_open`` / ``chan_put`` / ``chan_get``:
    a put to an actor of this process goes straight onto its queue, a put to a ``tcp://`` / ``ucx://`` endpoint goes
    through the transport, tensors bound for another process are parked in shared memory first.  Subclasses add where
    the hosted object runs (``construct`` / ``call`` / ``close``).
    """

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
            # pooled endpoint per peer, per-peer lock, one retry after eviction; CUDA tensors as IPC handles
            from ..transports import ucx as ucx_t

            host, port = tcp.parse_address(to_ep.address)
            await ucx_t.chan_put(host, port, to_ep.actor_id, name, payload)
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
        if ep.scheme == "ucx":
            from ..transports import ucx as ucx_t

            host, port = tcp.parse_address(ep.address)
            return await ucx_t.chan_get(host, port, ep.actor_id, name, timeout)
        if ep.scheme == "tcp":
            host, port = tcp.parse_address(ep.address)
            got = await tcp.chan_get(host, port, ep.actor_id, name, timeout)
            if isinstance(got, dict) and "__cuda_ipc__" in got:
                return cuda_ipc.loads(got["__cuda_ipc__"])
            return got
        raise RuntimeError("Endpoint mismatch")


__all__ = ["LocalMailboxBackend", "IN_HOST_SCHEMES", "IntraOpGovernor", "intra_op_governor"]
