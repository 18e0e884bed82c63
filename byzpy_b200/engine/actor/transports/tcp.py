"""One-shot request/reply helpers used by in-process backends to reach a TCP actor server's
mailboxes (reference engine/actor/transports/tcp.py:11-67)."""
from __future__ import annotations

import asyncio
from typing import Any, Optional, Tuple

from .._wire import recv_obj, send_obj


def parse_address(address: str) -> Tuple[str, int]:
    host, port = address.rsplit(":", 1)
    return host, int(port)


async def request(host: str, port: int, msg: dict, timeout: Optional[float] = None) -> Any:
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


async def chan_put(host: str, port: int, actor_id: str, name: str, payload: Any) -> None:
    await request(host, port, {"op": "chan_put", "actor_id": actor_id, "name": name, "payload": payload})


async def chan_get(host: str, port: int, actor_id: str, name: str, timeout: Optional[float]) -> Any:
    return await request(host, port, {"op": "chan_get", "actor_id": actor_id, "name": name,
                                      "timeout": timeout},
                         timeout=None if timeout is None else timeout + 5.0)


__all__ = ["chan_put", "chan_get", "request", "parse_address"]
