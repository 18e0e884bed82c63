"""Process actor: the object lives in a ``spawn``-ed child process behind a duplex pipe
(reference engine/actor/backends/process.py:19-321).

The class/factory travels by value (cloudpickle); tensor arguments and results cross the
boundary through POSIX shared memory (``wrap_payload`` / ``unwrap_payload``), one copy each way.
A lock serialises request/response pairs on the pipe; the blocking pipe I/O runs in the default
executor.  Mailboxes stay on the parent side (they are only read by the coordinator).
This is the CPU plumbing path; device-resident work uses the ``gpu`` backend instead.
"""
from __future__ import annotations

import asyncio
import multiprocessing as mp
import os
import threading
import traceback
import weakref
from typing import Any

import cloudpickle

from ..ipc import discard_payload, unwrap_payload, wrap_payload
from ._local import LocalMailboxBackend


def _child_main(conn) -> None:
    import asyncio as _aio
    import inspect as _inspect

    obj = None
    threads = None
    while True:
        try:
            msg = conn.recv()
        except (EOFError, OSError):
            break
        op = msg.get("op")
        try:
            want = msg.get("threads")
            if want and want != threads:            # this host's process actors split the cores between them
                import torch as _torch

                _torch.set_num_threads(int(want))
                threads = want
            if op == "stop":
                conn.send({"ok": True, "payload": None})
                break
            if op == "construct":
                target = cloudpickle.loads(msg["cls"])
                obj = target(*unwrap_payload(msg["args"]), **unwrap_payload(msg["kwargs"]))
                conn.send({"ok": True, "payload": None})
            elif op == "call":
                fn = getattr(obj, msg["method"])
                out = fn(*unwrap_payload(msg["args"]), **unwrap_payload(msg["kwargs"]))
                if _inspect.isawaitable(out):
                    out = _aio.run(_await(out))
                conn.send({"ok": True, "payload": wrap_payload(out)})
            else:
                conn.send({"ok": False, "error": f"unknown op {op!r}"})
        except BaseException as exc:  # noqa: BLE001 - everything goes back to the parent
            try:
                conn.send({"ok": False, "error": f"{exc!r}\n{traceback.format_exc()}"})
            except Exception:
                break
    try:
        conn.close()
    except Exception:
        pass


async def _await(x):
    return await x


_LIVE: "weakref.WeakSet[ProcessActorBackend]" = weakref.WeakSet()

_PRELOAD = ("numpy", "torch", "cloudpickle", "byzpy_b200")
_ctx_cache = {}


def process_context():
    """Multiprocessing context of every child this package starts.  Same semantics as ``spawn`` (a fresh
    interpreter state per child, ``__main__`` re-imported, nothing inherited by accident), but children are
    forked from a fork-server that has already imported torch and this package: a process actor is up in
    tens of milliseconds instead of the 3 s of a cold ``import torch``.  The server itself never touches
    CUDA or OpenMP (imports only), so forking from it is safe.  ``BYZPY_MP_START=spawn`` restores plain
    spawn (also the fallback where the platform has no fork-server)."""
    method = os.environ.get("BYZPY_MP_START", "forkserver")
    if method not in _ctx_cache:
        try:
            ctx = mp.get_context(method)
            if method == "forkserver":
                ctx.set_forkserver_preload(list(_PRELOAD))
        except ValueError:
            ctx = mp.get_context("spawn")
        _ctx_cache[method] = ctx
    return _ctx_cache[method]


def _thread_share() -> int:
    """Intra-op threads each live process actor of this parent may use: ``cores // actors``.  Five node
    processes each opening 8-thread OpenMP regions on 8 cores spend their time spinning at barriers (the
    reference's process examples show exactly that); the share travels with every request, so it follows
    actors being created and closed.  ``BYZPY_INTRAOP_GOVERNOR=0`` turns it off."""
    if os.environ.get("BYZPY_INTRAOP_GOVERNOR", "1") in ("0", "false", "False"):
        return 0
    live = sum(1 for b in _LIVE if not b._closed)
    return max(1, (os.cpu_count() or 1) // max(1, live))


class ProcessActorBackend(LocalMailboxBackend):
    """Hosts the object in a child process of its own (``"process"``).

    The class and every call travel over a duplex pipe (cloudpickle); tensors in arguments and results are parked in POSIX
    shared memory instead of being pickled, and released when the call is over -- also when it is cancelled.  Children
    are started with the ``forkserver`` method (``BYZPY_MP_START=spawn`` to change it) and share the host's cores: each
    call tells the child how many intra-op threads it may use.
    """

    scheme = "process"

    def __init__(self) -> None:
        super().__init__()
        self._closed = False
        _LIVE.add(self)
        ctx = process_context()
        self._conn, child = ctx.Pipe(duplex=True)
        self._proc = ctx.Process(target=_child_main, args=(child,), daemon=True)
        self._proc.start()
        child.close()
        self._io_lock = threading.Lock()
        self._closed = False

    async def start(self) -> None:
        if self._loop is None:
            self._loop = asyncio.get_running_loop()

    def _roundtrip(self, msg: dict) -> Any:
        with self._io_lock:
            if self._closed:
                raise RuntimeError("process actor is closed")
            try:
                share = _thread_share()
                if share:
                    msg = {**msg, "threads": share}
                self._conn.send(msg)
                reply = self._conn.recv()
            except (EOFError, OSError, BrokenPipeError) as exc:
                raise RuntimeError(f"process actor died (pipe error: {exc!r}); the child may have "
                                   "crashed or been closed concurrently") from exc
        if not reply.get("ok", False):
            raise RuntimeError(f"process actor error: {reply.get('error')}")
        return reply.get("payload")

    async def _send(self, msg: dict) -> Any:
        loop = asyncio.get_running_loop()
        fut = loop.run_in_executor(None, self._roundtrip, msg)
        try:
            return await asyncio.shield(fut)
        except asyncio.CancelledError:
            # the caller gave up (ParameterServer(node_timeout=...) wraps calls in wait_for); the round trip
            # still completes on its thread, and its reply may hold shared-memory segments nobody will unwrap
            def _discard(f) -> None:
                if not f.cancelled() and f.exception() is None:
                    try:
                        discard_payload(f.result())
                    except Exception:
                        pass

            fut.add_done_callback(_discard)
            raise

    async def construct(self, cls_or_factory: Any, *, args: tuple, kwargs: dict) -> None:
        await self._send({"op": "construct", "cls": cloudpickle.dumps(cls_or_factory),
                          "args": wrap_payload(tuple(args)), "kwargs": wrap_payload(dict(kwargs))})

    async def call(self, method: str, *args, **kwargs) -> Any:
        out = await self._send({"op": "call", "method": method, "args": wrap_payload(tuple(args)),
                                "kwargs": wrap_payload(dict(kwargs))})
        return unwrap_payload(out)

    async def close(self) -> None:
        if self._closed:
            return
        self._unregister()
        try:
            await self._send({"op": "stop"})
        except Exception:
            pass
        self._closed = True
        try:
            self._conn.close()
        except Exception:
            pass
        self._proc.join(timeout=5)
        if self._proc.is_alive():
            self._proc.terminate()


__all__ = ["ProcessActorBackend", "process_context"]
