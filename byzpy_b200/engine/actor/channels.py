"""Addresses and mailbox handles of the actor layer (reference engine/actor/channels.py:13-65).

An ``Endpoint`` names an actor globally: the transport scheme, where the hosting process listens
(empty for in-host schemes) and an id unique there.  A ``ChannelRef`` is a *local* actor's handle on
one of its named mailboxes: it posts to the same-named mailbox of any other endpoint and receives
from its own.  Payloads that crossed a process boundary arrive with tensors parked in shared
memory; ``recv`` materialises them.
"""
from __future__ import annotations

from typing import Any, NamedTuple, Optional

from .ipc import unwrap_payload


class Endpoint(NamedTuple):
    """Global address of an actor: ``(scheme, address, actor_id)``.

    ``scheme`` is the kind of backend hosting it (``"thread"``, ``"process"``, ``"gpu"``, ``"tcp"``, ``"ucx"``),
    ``address`` the ``host:port`` its server listens on (empty for in-host schemes), ``actor_id`` a unique id there.
    A plain named tuple: it pickles and travels inside messages, which is how actors introduce each other.

    Examples
    --------
    >>> from byzpy_b200.engine.actor.channels import Endpoint
    >>> ep = Endpoint("tcp", "10.0.0.5:29000", "abc")
    >>> str(ep), ep.is_remote()
    ('tcp://10.0.0.5:29000/abc', True)
    """

    scheme: str      # "thread" | "process" | "gpu" | "tcp" | "ucx"
    address: str     # "host:port" for tcp / ucx, "" otherwise
    actor_id: str

    def is_remote(self) -> bool:
        return self.scheme in ("tcp", "ucx")

    def __str__(self) -> str:
        where = f"//{self.address}" if self.address else ""
        return f"{self.scheme}:{where}/{self.actor_id}"


class ChannelRef:
    """A local actor's handle on one of its named mailboxes.

    ``await ch.send(endpoint, payload)`` posts to the same-named mailbox of the actor at ``endpoint`` (whatever backend
    hosts it -- delivery is in-process when possible, else shared memory, TCP or CUDA-IPC); ``await ch.recv(timeout=None)``
    takes the next payload from its own mailbox (``None`` on timeout); ``ch.endpoint`` is the address to give peers.
    Obtained from ``await ActorRef.open_channel(name)`` / :func:`open_channel`; a runnable tour is in
    ``examples/actor_demo/actor_demo.py``.
    """

    __slots__ = ("_backend", "_local", "_name")

    def __init__(self, backend, local_ep: Endpoint, name: str):
        self._backend, self._local, self._name = backend, local_ep, name

    def __repr__(self) -> str:
        return f"ChannelRef({self._name!r} @ {self._local})"

    @property
    def name(self) -> str:
        return self._name

    @property
    def endpoint(self) -> Endpoint:
        """Address peers use to reach this mailbox's owner."""
        return self._local

    async def send(self, to: Endpoint, payload: Any) -> None:
        """Post ``payload`` to mailbox ``name`` of the actor at ``to``."""
        await self._backend.chan_put(from_ep=self._local, to_ep=to, name=self._name, payload=payload)

    async def recv(self, *, timeout: Optional[float] = None) -> Any:
        """Next payload from this actor's own mailbox; ``None`` when ``timeout`` seconds pass without one (the
        reference's convention, backends/thread.py:127-137)."""
        item = await self._backend.chan_get(ep=self._local, name=self._name, timeout=timeout)
        return unwrap_payload(item)


async def open_channel(backend, name: str) -> ChannelRef:
    """Open (or attach to) mailbox ``name`` on the actor hosted by ``backend``."""
    endpoint = await backend.chan_open(name)
    return ChannelRef(backend, endpoint, name)


__all__ = ["Endpoint", "ChannelRef", "open_channel"]
