"""One-shot request/reply helpers used by in-process backends to reach a TCP actor server's
mailboxes (reference engine/actor/transports/tcp.py:11-67)."""
from __future__ import annotations

import asyncio
from typing import Any, Optional, Tuple

from .._wire import recv_obj, send_obj


def parse_address(address: str) -> Tuple[str, int]:
    """``"host:port"`` -> ``(host, port)``."""
    host, port = address.rsplit(":", 1)
    return host, int(port)


async def request(host: str, port: int, msg: dict, timeout: Optional[float] = None) -> Any:
    """One request/response exchange with the actor server at ``host:port`` over a fresh connection; raises what the
    server reports.
    """
    reader, writer = await asyncio.open_connection(host, port)
    try:
        await send_obj(writer, msg)
        reply = await asyncio.wait_for(recv_obj(reader), timeout=timeout)
    finally:
        writer.close()
        try:
            await writer.wait_closed()
        except Exception:
            pass
    if isinstance(reply, dict) and reply.get("ok") is False:
        raise RuntimeError(reply.get("error", "remote error"))
    return reply.get("payload") if isinstance(reply, dict) else reply


def _target(address, port, actor_id, to_ep):
    """(host, port, actor id) from either calling convention: ``(host, port, actor_id, ...)`` as used inside this
    package, or the reference's ``(address, *, to_ep=..., ...)`` with ``address = "host:port"`` (reference
    transports/tcp.py:27-67)."""
    if port is None:
        host, port = parse_address(address)
    else:
        host = address
    if actor_id is None and to_ep is not None:
        actor_id = to_ep["actor_id"] if isinstance(to_ep, dict) else getattr(to_ep, "actor_id")
    if actor_id is None:
        raise TypeError("an actor id (or to_ep) is required")
    return host, int(port), actor_id


async def chan_put(address: str, port: Optional[int] = None, actor_id: Optional[str] = None, name: Optional[str] = None,
                   payload: Any = None, *, from_ep: Any = None, to_ep: Any = None) -> None:
    """Post ``payload`` to mailbox ``name`` of a remote actor.  Accepts ``(host, port, actor_id, name, payload)`` and the
    reference's ``(address, to_ep=..., name=..., payload=...)`` calling convention.
    """
    host, port, actor_id = _target(address, port, actor_id, to_ep)
    msg = {"op": "chan_put", "actor_id": actor_id, "name": name, "payload": payload}
    if to_ep is not None and not isinstance(to_ep, dict):
        msg["to"] = (to_ep.scheme, to_ep.address, to_ep.actor_id)
    await request(host, port, msg)


async def chan_get(address: str, port: Optional[int] = None, actor_id: Optional[str] = None, name: Optional[str] = None,
                   timeout: Optional[float] = None) -> Any:
    """Take the next payload from mailbox ``name`` of a remote actor (``None`` on timeout); both calling conventions of
    :func:`chan_put`.
    """
    host, port, actor_id = _target(address, port, actor_id, None)
    return await request(host, port, {"op": "chan_get", "actor_id": actor_id, "name": name,
                                      "timeout": timeout},
                         timeout=None if timeout is None else timeout + 5.0)


__all__ = ["chan_put", "chan_get", "request", "parse_address"]
