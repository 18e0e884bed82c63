"""TCP actor server and its client backend (reference engine/actor/backends/remote.py:19-433).

Protocol: one persistent connection per client backend; each request is a framed dict
``{"op": ..., ...}`` answered by ``{"ok": bool, "payload"|"error": ...}``.  Ops: ``construct``,
``call``, ``get_ep``, ``chan_open``, ``chan_put``, ``chan_get``, ``close``.  The server keeps
``actor_id -> object`` and per-actor mailboxes; synchronous methods run in the default executor.
Tensors travel by value (valid across hosts); with ``gpu_direct`` (the ``ucx`` scheme) CUDA
tensors travel as CUDA IPC handles on the same host.
"""
from __future__ import annotations

import asyncio
import inspect
import traceback
import uuid
from typing import Any, Dict, Optional

import cloudpickle

from .._wire import recv_obj, send_obj
from ..channels import Endpoint
from ..transports import cuda_ipc


def _is_local(host: str) -> bool:
    return host in ("127.0.0.1", "localhost", "0.0.0.0", "::1")


class RemoteActorBackend:
    """Client side of an actor hosted by a :class:`RemoteActorServer` on another machine (``"tcp://host:port"``).

    One TCP connection per backend; ``construct`` ships the class by value, calls are request/response frames
    (length-prefixed pickles) serialised by a lock.  If a call is cancelled or fails mid-exchange the connection is
    dropped and re-opened on the next call, so a late reply can never be mistaken for the answer to a later request.

    Parameters
    ----------
    host, port :
        Address of the server.
    """

    scheme = "tcp"
    gpu_direct = False

    def __init__(self, host: str, port: int) -> None:
        self.host, self.port = host, int(port)
        self._reader: Optional[asyncio.StreamReader] = None
        self._writer: Optional[asyncio.StreamWriter] = None
        self._lock = asyncio.Lock()
        self._actor_id: Optional[str] = None

    @property
    def _address(self) -> str:
        return f"{self.host}:{self.port}"

    def _pack(self, obj: Any) -> Any:
        if self.gpu_direct:
            return {"__cuda_ipc__": cuda_ipc.dumps(obj, same_host=_is_local(self.host))}
        return cuda_ipc._to_host(obj)

    @staticmethod
    def _unpack(obj: Any) -> Any:
        if isinstance(obj, dict) and "__cuda_ipc__" in obj:
            return cuda_ipc.loads(obj["__cuda_ipc__"])
        return obj

    async def start(self) -> None:
        if self._writer is None:
            self._reader, self._writer = await asyncio.open_connection(self.host, self.port)

    def _drop_connection(self) -> None:
        writer, self._reader, self._writer = self._writer, None, None
        if writer is not None:
            try:
                writer.close()
            except Exception:
                pass

    async def _rpc(self, msg: Dict[str, Any]) -> Any:
        async with self._lock:
            await self.start()
            try:
                await send_obj(self._writer, msg)
                reply = await recv_obj(self._reader)
            except BaseException:
                # cancelled (a caller's wait_for timed out -- ParameterServer(node_timeout=...) does that) or
                # broken in mid-exchange: the reply may still arrive, and the next exchange on this connection
                # would take it for its own.  The actor lives on in the server under its id; reconnect next time.
                self._drop_connection()
                raise
        if not reply.get("ok", False):
            raise RuntimeError(f"remote actor error: {reply.get('error')}")
        return reply.get("payload")

    async def construct(self, cls_or_factory: Any, *, args: tuple, kwargs: dict) -> None:
        blob = cloudpickle.dumps(cls_or_factory)
        self._actor_id = await self._rpc({"op": "construct", "cls": blob,
                                          "args": self._pack(tuple(args)),
                                          "kwargs": self._pack(dict(kwargs))})

    async def call(self, method: str, *args, **kwargs) -> Any:
        if self._actor_id is None:
            raise RuntimeError("actor not constructed")
        out = await self._rpc({"op": "call", "actor_id": self._actor_id, "method": method,
                               "args": self._pack(tuple(args)), "kwargs": self._pack(dict(kwargs))})
        return self._unpack(out)

    async def get_endpoint(self) -> Endpoint:
        if self._actor_id is None:
            raise RuntimeError("actor not constructed")
        return Endpoint(scheme=self.scheme, address=self._address, actor_id=self._actor_id)

    async def chan_open(self, name: str) -> Endpoint:
        await self._rpc({"op": "chan_open", "actor_id": self._actor_id, "name": name})
        return await self.get_endpoint()

    async def chan_put(self, *, from_ep: Endpoint, to_ep: Endpoint, name: str, payload: Any) -> None:
        if to_ep.scheme in ("thread", "process", "gpu"):
            # the target lives in THIS (client) process: deliver through the local router
            from ..router import channel_router

            peer = channel_router.resolve(to_ep.scheme, to_ep.actor_id)
            if peer is None:
                raise RuntimeError(f"no local {to_ep.scheme} actor {to_ep.actor_id}")
            await peer._deliver_local(name, from_ep, payload)
            return
        await self._rpc({"op": "chan_put", "actor_id": to_ep.actor_id, "name": name,
                         "to": (to_ep.scheme, to_ep.address, to_ep.actor_id),
                         "payload": self._pack(payload)})

    async def chan_get(self, *, ep: Endpoint, name: str, timeout: Optional[float]) -> Any:
        out = await self._rpc({"op": "chan_get", "actor_id": ep.actor_id, "name": name,
                               "timeout": timeout})
        return self._unpack(out)

    async def close(self) -> None:
        if self._writer is None:
            return
        try:
            if self._actor_id is not None:
                await self._rpc({"op": "close", "actor_id": self._actor_id})
        except Exception:
            pass
        try:
            self._writer.close()
            await self._writer.wait_closed()
        except Exception:
            pass
        self._reader = self._writer = None


class RemoteActorServer:
    """Hosts actors for remote clients: ``await RemoteActorServer(host, port).serve()``.

    Each client connection may construct actors, call them, open their mailboxes and post to them.  Frames are
    unpickled, so whoever can reach the port can run code in this process: the default binds the loopback interface;
    pass ``host="0.0.0.0"`` explicitly on a trusted network.

    Parameters
    ----------
    host : str, default "127.0.0.1"
    port : int, default 29000
        0 picks a free port (``port`` holds it after ``start()``).

    Notes
    -----
    ``await start()`` binds without blocking, ``await serve()`` binds and serves until cancelled, ``await stop()``
    closes the listener.  :func:`start_actor_server` is ``serve()`` as a one-liner for scripts
    (``examples/distributed/server.py``).
    """

    scheme = "tcp"
    gpu_direct = False

    def __init__(self, host: str = "127.0.0.1", port: int = 29000) -> None:
        # Frames are unpickled: whoever can reach the port can run code in this process.  The default
        # therefore binds the loopback interface; pass the interface to expose explicitly
        # (``host="0.0.0.0"``) on a trusted network only -- the reference has no default and its
        # examples bind 0.0.0.0 (reference engine/actor/backends/remote.py:264-290).
        self.host, self.port = host, int(port)
        self._actors: Dict[str, Any] = {}
        self._mailboxes: Dict[str, Dict[str, asyncio.Queue]] = {}
        self._server: Optional[asyncio.AbstractServer] = None

    # ------------------------------------------------------------------ helpers
    def _box(self, actor_id: str, name: str) -> asyncio.Queue:
        return self._mailboxes.setdefault(actor_id, {}).setdefault(name, asyncio.Queue())

    @staticmethod
    def _unpack(obj: Any) -> Any:
        if isinstance(obj, dict) and "__cuda_ipc__" in obj:
            return cuda_ipc.loads(obj["__cuda_ipc__"])
        return obj

    def _pack(self, obj: Any) -> Any:
        if self.gpu_direct:
            return {"__cuda_ipc__": cuda_ipc.dumps(obj, same_host=True)}
        return cuda_ipc._to_host(obj)

    async def _invoke(self, fn, args, kwargs):
        if inspect.iscoroutinefunction(fn):
            return await fn(*args, **kwargs)
        loop = asyncio.get_running_loop()
        return await loop.run_in_executor(None, lambda: fn(*args, **kwargs))

    async def _dispatch(self, msg: Dict[str, Any]) -> Any:
        op = msg.get("op")
        if op == "construct":
            target = cloudpickle.loads(msg["cls"])
            obj = await self._invoke(target, self._unpack(msg["args"]), self._unpack(msg["kwargs"]))
            actor_id = str(uuid.uuid4())
            self._actors[actor_id] = obj
            return actor_id
        if op == "call":
            obj = self._actors[msg["actor_id"]]
            out = await self._invoke(getattr(obj, msg["method"]), self._unpack(msg["args"]),
                                     self._unpack(msg["kwargs"]))
            return self._pack(out)
        if op == "get_ep":
            return (self.scheme, f"{self.host}:{self.port}", msg["actor_id"])
        if op == "chan_open":
            self._box(msg["actor_id"], msg["name"])
            return True
        if op == "chan_put":
            to = msg.get("to")
            if to is not None and to[1] and to[1] != f"{self.host}:{self.port}" and to[0] in ("tcp", "ucx"):
                # relay to another server
                from ..transports import tcp as tcp_t

                h, p = tcp_t.parse_address(to[1])
                await tcp_t.chan_put(h, p, to[2], msg["name"], msg["payload"])
                return True
            await self._box(msg["actor_id"], msg["name"]).put(msg["payload"])
            return True
        if op == "chan_get":
            q = self._box(msg["actor_id"], msg["name"])
            timeout = msg.get("timeout")
            try:
                item = await (q.get() if timeout is None else asyncio.wait_for(q.get(), timeout))
            except asyncio.TimeoutError:
                return None
            return item
        if op == "close":
            self._actors.pop(msg.get("actor_id"), None)
            self._mailboxes.pop(msg.get("actor_id"), None)
            return True
        raise ValueError(f"unknown op {op!r}")

    async def _handle(self, reader: asyncio.StreamReader, writer: asyncio.StreamWriter) -> None:
        try:
            while True:
                try:
                    msg = await recv_obj(reader)
                except (asyncio.IncompleteReadError, ConnectionError):
                    break
                try:
                    reply = {"ok": True, "payload": await self._dispatch(msg)}
                except Exception as exc:  # error travels back to the caller
                    reply = {"ok": False, "error": f"{exc!r}\n{traceback.format_exc()}"}
                await send_obj(writer, reply)
        finally:
            try:
                writer.close()
            except Exception:
                pass

    async def start(self) -> None:
        if self._server is None:
            self._server = await asyncio.start_server(self._handle, self.host, self.port)
            sock = self._server.sockets[0].getsockname()
            self.port = sock[1]

    async def serve(self) -> None:
        await self.start()
        print(f"[byzpy_b200] {self.scheme} actor server listening on {self.host}:{self.port}", flush=True)
        async with self._server:
            await self._server.serve_forever()

    async def stop(self) -> None:
        if self._server is not None:
            self._server.close()
            await self._server.wait_closed()
            self._server = None


async def start_actor_server(host: str, port: int) -> None:
    """Serve a :class:`RemoteActorServer` on ``host:port`` until cancelled."""
    await RemoteActorServer(host, port).serve()


__all__ = ["RemoteActorBackend", "RemoteActorServer", "start_actor_server"]
