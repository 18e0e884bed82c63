"""Data plane of the ``ucx://`` scheme on one NVSwitch box.

The reference moves device tensors between actor servers through UCX endpoints: a pooled endpoint
per ``(host, port)``, one lock per peer so a control message + payload + reply exchange is never
interleaved, and one retry on a transport error after evicting the endpoint
(reference engine/actor/transports/ucx.py:110-133), with CUDA payloads sent as raw device buffers
behind a pickled descriptor (``:186-277``).  On a single B200 node no copy engine is needed in the
middle: every GPU reaches every other at full NVLink bandwidth, so the payload of a CUDA tensor is
its 64-byte CUDA-IPC handle (``transports/cuda_ipc.py``) and the receiver maps the sender's memory
-- zero copy, device to device.  What remains of the transport is the connection management, which
this module provides with the reference's surface:

* :func:`get_endpoint` / :func:`call` / :func:`clear_pool` -- pooled stream endpoints per peer,
  per-peer locks, one retry after eviction on a broken connection;
* :func:`pack_payload` / :func:`unpack_payload` -- ``("cuda" | "pickle", blob)`` payload tagging;
* :func:`chan_put` / :func:`chan_get` -- mailbox access of a remote ``ucx://`` actor server.

A ``ucx://`` endpoint on ANOTHER host gets by-value payloads (CUDA IPC handles are only valid on the
box that exported them); ``have_ucx()`` tells whether the GPU-direct path is usable at all.
"""
from __future__ import annotations

import asyncio
from typing import Any, Awaitable, Callable, Dict, Optional, Tuple

from .._wire import recv_obj, send_obj
from . import cuda_ipc

Endpoint = Tuple[asyncio.StreamReader, asyncio.StreamWriter]

_EP_CACHE: Dict[Tuple[int, str, int], Endpoint] = {}
_EP_LOCKS: Dict[Tuple[int, str, int], asyncio.Lock] = {}
_RETRYABLE = (ConnectionError, asyncio.IncompleteReadError, BrokenPipeError, EOFError, OSError)


def _ucx_mod():
    """The installed UCX Python binding (``ucxx``, else ``ucp``), or None -- the probe the reference's
    callers use (reference engine/actor/transports/ucx.py:36-56).  Nothing here needs it."""
    import importlib

    for name in ("ucxx", "ucp"):
        try:
            return importlib.import_module(name)
        except Exception:  # noqa: BLE001  (absent, or present but unusable on this box)
            continue
    return None


def have_ucx() -> bool:
    """True when the GPU-direct payload path is usable (a CUDA device is present)."""
    return cuda_ipc.available()


def _key(host: str, port: int) -> Tuple[int, str, int]:
    """Pool key: streams and locks belong to the event loop that created them, so a process that runs several
    loops one after the other (``asyncio.run`` per call) gets one endpoint per (loop, peer)."""
    try:
        loop_id = id(asyncio.get_running_loop())
    except RuntimeError:
        loop_id = 0
    return (loop_id, host, int(port))


def _drop_dead_loops() -> None:
    """Forget endpoints whose event loop is gone (their transports were closed with the loop)."""
    for key in [k for k, (_, w) in _EP_CACHE.items() if w.is_closing()]:
        _EP_CACHE.pop(key, None)
        _EP_LOCKS.pop(key, None)


def is_same_host(host: str) -> bool:
    """True for loopback addresses: CUDA-IPC handles are only meaningful on the machine that exported them."""
    return host in ("127.0.0.1", "localhost", "0.0.0.0", "::1")


async def get_endpoint(host: str, port: int) -> Endpoint:
    """The pooled connection to ``host:port`` (opened on first use, re-opened after eviction)."""
    key = _key(host, port)
    ep = _EP_CACHE.get(key)
    if ep is not None and not ep[1].is_closing():
        return ep
    if len(_EP_CACHE) > 64:
        _drop_dead_loops()
    reader, writer = await asyncio.open_connection(host, int(port))
    _EP_CACHE[key] = (reader, writer)
    return reader, writer


def evict_endpoint(host: str, port: int) -> None:
    """Close and forget the pooled connection to ``host:port`` (the next use re-opens it)."""
    ep = _EP_CACHE.pop(_key(host, port), None)
    if ep is not None:
        try:
            ep[1].close()
        except Exception:
            pass


async def clear_pool() -> None:
    """Close every pooled connection of this process."""
    here = _key("", 0)[0]
    for key in list(_EP_CACHE):
        _, writer = _EP_CACHE.pop(key)
        try:
            writer.close()
            if key[0] == here:              # (a stream of another loop cannot be awaited from this one)
                await writer.wait_closed()
        except Exception:
            pass
    _EP_LOCKS.clear()


async def call(host: str, port: int, fn: Callable[[Endpoint], Awaitable[Any]]) -> Any:
    """Run one full exchange ``fn(endpoint)`` on the pooled endpoint of the peer, serialised by the
    peer's lock; a broken connection evicts the endpoint and the exchange is retried once."""
    key = _key(host, port)
    lock = _EP_LOCKS.setdefault(key, asyncio.Lock())

    async def _once():
        async with lock:
            ep = await get_endpoint(host, port)
            return await fn(ep)

    for attempt in (0, 1):
        try:
            return await _once()
        except (asyncio.TimeoutError, asyncio.CancelledError):
            # the peer's reply is still in flight: the next exchange on this connection would read it as ITS
            # reply, so the endpoint is not reusable (and a timeout is the caller's answer, not a broken link)
            evict_endpoint(host, port)
            raise
        except _RETRYABLE:
            evict_endpoint(host, port)
            if attempt == 1:
                raise
        except BaseException:
            evict_endpoint(host, port)
            raise


async def create_endpoint(host: str, port: int) -> Endpoint:
    """A dedicated endpoint that bypasses the pool -- for a persistent one-to-one connection with a server
    (reference transports/ucx.py:84-93)."""
    return await asyncio.open_connection(host, int(port))


async def create_listener(cb: Callable[[Endpoint], Any], *, host: str, port: int):
    """Listen on ``host:port``; ``cb(endpoint)`` (plain function or coroutine function) is invoked per accepted
    connection with its ``(reader, writer)`` endpoint and the connection is closed when it returns (reference
    transports/ucx.py:136-160 wraps
    ``ucxx.create_listener``).  Returns the ``asyncio`` server; ``port=0`` picks a free port
    (``server.sockets[0].getsockname()[1]``)."""

    async def _accept(reader: asyncio.StreamReader, writer: asyncio.StreamWriter) -> None:
        try:
            out = cb((reader, writer))
            if asyncio.iscoroutine(out):
                await out
        except (asyncio.IncompleteReadError, ConnectionError):
            pass                                   # the peer hung up in the middle of an exchange
        finally:
            writer.close()                         # the connection lives as long as its handler

    return await asyncio.start_server(_accept, host, int(port))


async def send_control(ep: Endpoint, obj: Dict[str, Any]) -> None:
    """One length-prefixed control message on an endpoint (reference transports/ucx.py:210-214)."""
    await send_obj(ep[1], obj)


async def recv_control(ep: Endpoint) -> Dict[str, Any]:
    """Read one control frame (a dict) from the endpoint."""
    return await recv_obj(ep[0])


async def send_payload(ep: Endpoint, tag: str, desc: Any, obj: Any = None) -> None:
    """Send what :func:`pack_payload` produced: ``tag`` is ``"cuda"`` (``desc`` = CUDA-IPC handles: the receiver
    maps the sender's device memory, nothing else travels) or ``"pickle"`` (``desc`` = the by-value blob).  The
    reference sends the device buffer itself after the descriptor (transports/ucx.py:245-254); on one NVSwitch box
    the handle IS the transfer."""
    if not isinstance(desc, (bytes, bytearray)):
        tag, desc = pack_payload(obj if obj is not None else desc)
    await send_control(ep, {"ptag": tag, "desc": bytes(desc)})


async def recv_payload(ep: Endpoint) -> Any:
    """Read one payload written by :func:`send_payload` and rebuild it (CUDA tensors arrive as mapped device memory)."""
    ctrl = await recv_control(ep)
    return unpack_payload(ctrl["desc"])


def pack_payload(obj: Any, *, same_host: bool = True) -> Tuple[str, bytes]:
    """``(tag, bytes)`` of a payload: CUDA-IPC handles when ``same_host``, host copies otherwise."""
    blob = cuda_ipc.dumps(obj, same_host=same_host)
    return ("cuda" if blob[:1] == b"I" else "pickle"), blob


def unpack_payload(blob: bytes) -> Any:
    """Inverse of :func:`pack_payload`."""
    return cuda_ipc.loads(blob)


async def request(host: str, port: int, msg: dict, timeout: Optional[float] = None) -> Any:
    """One control exchange with a ``ucx://`` actor server over the pooled endpoint."""

    async def _do(ep: Endpoint):
        reader, writer = ep
        await send_obj(writer, msg)
        return await asyncio.wait_for(recv_obj(reader), timeout=timeout)

    reply = await call(host, port, _do)
    if isinstance(reply, dict) and reply.get("ok") is False:
        raise RuntimeError(reply.get("error", "remote error"))
    return reply.get("payload") if isinstance(reply, dict) else reply


async def chan_put(host: str, port: int, actor_id: str, name: str, payload: Any) -> None:
    """Deliver ``payload`` into mailbox ``name`` of a remote ``ucx://`` actor; CUDA tensors travel as
    CUDA-IPC handles when the server is on this box."""
    _, blob = pack_payload(payload, same_host=is_same_host(host))
    await request(host, port, {"op": "chan_put", "actor_id": actor_id, "name": name,
                               "payload": {"__cuda_ipc__": blob}})


async def chan_get(host: str, port: int, actor_id: str, name: str, timeout: Optional[float]) -> Any:
    """Take the next payload from mailbox ``name`` of actor ``actor_id`` hosted at ``host:port`` (``None`` on timeout)."""
    out = await request(host, port, {"op": "chan_get", "actor_id": actor_id, "name": name, "timeout": timeout},
                        timeout=None if timeout is None else timeout + 5.0)
    if isinstance(out, dict) and "__cuda_ipc__" in out:
        return unpack_payload(out["__cuda_ipc__"])
    return out


__all__ = ["have_ucx", "pack_payload", "unpack_payload", "get_endpoint", "create_endpoint", "create_listener",
           "evict_endpoint", "clear_pool", "call", "request", "send_control", "recv_control", "send_payload",
           "recv_payload", "chan_put", "chan_get", "is_same_host"]
