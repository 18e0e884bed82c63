"""GPU actor backend: a CUDA-stream worker.

The reference's ``GPUActorBackend`` is an in-process asyncio actor with no CUDA code
(reference engine/actor/backends/gpu.py:23-200).  Here a ``"gpu"`` actor owns a device and a
dedicated CUDA stream: every method call runs on a private host thread with that stream current,
so all kernels the actor launches (the hand-written sm_100a kernels and PyTorch ops alike) are
issued on the actor's own stream and several gpu actors overlap ON THE DEVICE.  Ordering across
actors is device-side: the actor stream first waits on the caller's current stream, and the
caller's stream waits on an event recorded after the call -- no host synchronisation on the
path.  Same-process gpu->gpu channel sends enqueue the tensor object itself (zero copy).
Without CUDA the backend degrades to a plain thread actor so CPU-only tests still run.

``UCXRemoteActorBackend`` / ``UCXRemoteActorServer`` / ``start_ucx_actor_server`` keep the
reference names (reference gpu.py:206-662) for the ``ucx://host:port`` scheme; on a single
NVSwitch box the GPU-direct payload path is CUDA IPC (see ``transports/cuda_ipc.py``) riding
on the TCP control plane, instead of UCX tagged sends.
"""
from __future__ import annotations

import asyncio
import concurrent.futures
import inspect
import itertools
from typing import Any, Optional

from ._local import LocalMailboxBackend
from .remote import RemoteActorBackend, RemoteActorServer

_device_rr = itertools.count()


def _cuda_ok() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # pragma: no cover
        return False


def _record_stream(obj: Any, stream) -> None:
    """``tensor.record_stream(stream)`` for every CUDA tensor in a nested result."""
    if hasattr(obj, "is_cuda") and hasattr(obj, "record_stream"):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, (list, tuple)):
        for x in obj:
            _record_stream(x, stream)
    elif isinstance(obj, dict):
        for v in obj.values():
            _record_stream(v, stream)


class GPUActorBackend(LocalMailboxBackend):
    """Hosts the object on a worker thread bound to one CUDA device and stream (``"gpu"`` / ``"gpu:<index>"``).

    Every call runs with the actor's stream current, so kernels launched by different GPU actors overlap on the
    device; results are handed to the caller ordered by an event on that stream (and ``record_stream`` keeps the
    allocator from recycling them early) rather than by a device synchronisation.  Without CUDA it degrades to a thread
    actor.

    Parameters
    ----------
    device : int, optional
        CUDA device index; default: round-robin over the visible devices.
    """

    scheme = "gpu"

    def __init__(self, device: Optional[int] = None) -> None:
        super().__init__()
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=1,
                                                           thread_name_prefix="byz-gpu-actor")
        self._obj: Any = None
        self._device = None
        self._stream = None
        if _cuda_ok():
            import torch

            idx = device if device is not None else next(_device_rr) % torch.cuda.device_count()
            self._device = torch.device("cuda", idx)
            self._stream = torch.cuda.Stream(device=self._device)

    @property
    def device(self):
        return self._device

    @property
    def stream(self):
        return self._stream

    async def start(self) -> None:
        if self._loop is None:
            self._loop = asyncio.get_running_loop()

    async def _on_stream(self, fn, *args, **kwargs):
        loop = asyncio.get_running_loop()
        if self._stream is None:
            return await loop.run_in_executor(self._pool, lambda: fn(*args, **kwargs))
        import torch

        caller = torch.cuda.current_stream(self._device)
        ready = torch.cuda.Event()
        ready.record(caller)

        def _run():
            with torch.cuda.device(self._device), torch.cuda.stream(self._stream):
                self._stream.wait_event(ready)        # inputs produced on the caller's stream
                out = fn(*args, **kwargs)
                done = torch.cuda.Event()
                done.record(self._stream)
                return out, done

        out, done = await loop.run_in_executor(self._pool, _run)
        caller = torch.cuda.current_stream(self._device)
        caller.wait_event(done)                                    # device-side join
        # results were allocated on the ACTOR's stream and are about to be consumed on the caller's: tell the
        # caching allocator, or a block freed by the caller could be handed to this actor's next call (by a
        # different caller) while kernels of this caller still read it
        _record_stream(out, caller)
        return out

    async def construct(self, cls_or_factory: Any, *, args: tuple, kwargs: dict) -> None:
        self._obj = await self._on_stream(cls_or_factory, *args, **kwargs)

    async def call(self, method: str, *args, **kwargs) -> Any:
        if self._obj is None:
            raise RuntimeError("actor not constructed")
        fn = getattr(self._obj, method)
        if inspect.iscoroutinefunction(fn):
            return await fn(*args, **kwargs)
        return await self._on_stream(fn, *args, **kwargs)

    async def close(self) -> None:
        self._unregister()
        self._pool.shutdown(wait=True)


class UCXRemoteActorBackend(RemoteActorBackend):
    """Client of a ``ucx://host:port`` actor server: TCP control plane, CUDA-IPC payloads."""

    scheme = "ucx"
    gpu_direct = True


class UCXRemoteActorServer(RemoteActorServer):
    """Actor server for ``ucx://host:port`` clients: like :class:`~byzpy_b200.engine.actor.backends.remote.RemoteActorServer`,
    but CUDA tensors in calls, results and channel messages cross the process boundary as CUDA-IPC handles (the
    receiver maps the sender's memory; same machine only) instead of being staged through the host.

    Parameters
    ----------
    host : str, default "127.0.0.1"
    port : int, default 0
        0 picks a free port; ``address()`` tells which.
    """

    scheme = "ucx"
    gpu_direct = True

    def __init__(self, host: str = "127.0.0.1", port: int = 0) -> None:
        # port 0 = an ephemeral port, as in the reference (backends/gpu.py:470-480); the host default stays the
        # loopback interface (the reference listens on 0.0.0.0 and unpickles whatever connects)
        super().__init__(host, port)

    def address(self) -> str:
        """``host:port`` the server listens on (reference backends/gpu.py:490-491)."""
        return f"{self.host}:{self.port}"


async def start_ucx_actor_server(host: str = "127.0.0.1", port: int = 0) -> None:
    """Serve a :class:`UCXRemoteActorServer` on ``host:port`` until cancelled."""
    await UCXRemoteActorServer(host, port).serve()


__all__ = ["GPUActorBackend", "UCXRemoteActorBackend", "UCXRemoteActorServer",
           "start_ucx_actor_server"]
