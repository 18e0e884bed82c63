"""Heterogeneous worker pool executing ``SubTask`` s (reference engine/graph/pool.py:28-374).

Each worker is an actor (any backend) hosting a ``_SubTaskWorker``.  Workers carry capability
tags: ``"gpu"`` for ``"gpu"`` / ``ucx://`` backends, else ``"cpu"``, plus a unique
``worker::<name>-<idx>`` tag used for round-robin pinning.  ``run_subtask`` acquires an idle
worker with the requested affinity (parking a waiter when all matching workers are busy, raising
``RuntimeError("No actor in the pool")`` when none can ever match), honours ``max_retries`` and
releases the worker to capability-specific waiters first.  Functions are shipped by value
(cloudpickle) with a per-worker LRU of serialised payloads AND a worker-side LRU of
deserialised callables (the reference re-unpickles on every call).
"""
from __future__ import annotations

import asyncio
from collections import OrderedDict, defaultdict, deque
from dataclasses import dataclass
from typing import Any, Callable, Deque, Dict, List, Mapping, Optional, Sequence, Union

import cloudpickle

from ...utils import metrics
from ..actor.base import ActorBackend, ActorRef
from ..actor.channels import ChannelRef, Endpoint
from ..actor.factory import resolve_backend
from .subtask import SubTask


def _infer_capabilities(spec: Union[str, ActorBackend]) -> Sequence[str]:
    if isinstance(spec, str):
        return ("gpu",) if (spec == "gpu" or spec.startswith("gpu:") or spec.startswith("ucx://")) else ("cpu",)
    from ..actor.backends.gpu import GPUActorBackend, UCXRemoteActorBackend

    return ("gpu",) if isinstance(spec, (GPUActorBackend, UCXRemoteActorBackend)) else ("cpu",)


@dataclass(frozen=True)
class ActorPoolConfig:
    """One homogeneous group of workers of an :class:`ActorPool`.

    Parameters
    ----------
    backend : str or ActorBackend
        ``"thread"``, ``"process"``, ``"gpu"`` / ``"gpu:<index>"`` (a CUDA-stream worker), ``"tcp://host:port"``,
        ``"ucx://host:port"``, or a backend instance.
    count : int, default 1
        Number of workers of this kind.
    capabilities : sequence of str, optional
        Tags subtasks can ask for through ``SubTask.affinity``; default ``("gpu",)`` for GPU / ucx backends, else
        ``("cpu",)``.
    name : str, optional
        Prefix of the worker labels (``<name>-<index>``; default ``actor``).
    """

    backend: Union[str, ActorBackend]
    count: int = 1
    capabilities: Optional[Sequence[str]] = None
    name: Optional[str] = None

    def resolved_capabilities(self) -> Sequence[str]:
        return self.capabilities if self.capabilities is not None else _infer_capabilities(self.backend)


class _SubTaskWorker:
    """Lives inside the worker actor; runs pickled callables."""

    _CACHE_LIMIT = 64

    def __init__(self) -> None:
        self._fns: "OrderedDict[bytes, Callable[..., Any]]" = OrderedDict()

    def execute(self, payload: bytes, args: tuple, kwargs: Mapping[str, Any]) -> Any:
        fn = self._fns.get(payload)
        if fn is None:
            fn = cloudpickle.loads(payload)
            self._fns[payload] = fn
            if len(self._fns) > self._CACHE_LIMIT:
                self._fns.popitem(last=False)
        else:
            self._fns.move_to_end(payload)
        return fn(*args, **dict(kwargs))


class _PoolWorker:
    def __init__(self, *, backend: ActorBackend, capabilities: set, name: str) -> None:
        self.backend = backend
        self.capabilities = frozenset(capabilities)
        self.name = name
        self._ref = ActorRef(backend)
        self._endpoint: Optional[Endpoint] = None
        self._channels: Dict[str, ChannelRef] = {}
        self._fn_cache: "OrderedDict[Callable[..., Any], bytes]" = OrderedDict()
        self._fn_cache_limit = 64

    async def start(self) -> None:
        await self.backend.start()
        await self.backend.construct(_SubTaskWorker, args=(), kwargs={})

    async def endpoint(self) -> Endpoint:
        if self._endpoint is None:
            self._endpoint = await self.backend.get_endpoint()
        return self._endpoint

    async def open_channel(self, name: str) -> ChannelRef:
        if name not in self._channels:
            self._channels[name] = ChannelRef(self.backend, await self.backend.chan_open(name), name)
        return self._channels[name]

    def _serialized_fn(self, fn: Callable[..., Any]) -> bytes:
        try:
            blob = self._fn_cache.pop(fn)
        except (KeyError, TypeError):
            blob = cloudpickle.dumps(fn)
        try:
            self._fn_cache[fn] = blob
            if len(self._fn_cache) > self._fn_cache_limit:
                self._fn_cache.popitem(last=False)
        except TypeError:  # unhashable callable
            pass
        return blob

    async def run(self, subtask: SubTask) -> Any:
        return await self._ref.execute(self._serialized_fn(subtask.fn), tuple(subtask.args),
                                       dict(subtask.kwargs))

    async def close(self) -> None:
        await self.backend.close()


class ActorPoolChannel:
    """A channel bound on every pool worker; send/recv addressed by worker name."""

    def __init__(self, *, name: str, channels: Mapping[str, ChannelRef],
                 endpoints: Mapping[str, Endpoint]) -> None:
        self.name = name
        self._channels = dict(channels)
        self._endpoints = dict(endpoints)

    @property
    def workers(self) -> Sequence[str]:
        return tuple(self._channels)

    def channel(self, worker: str) -> ChannelRef:
        try:
            return self._channels[worker]
        except KeyError as exc:
            raise KeyError(f"No channel bound for worker {worker!r}") from exc

    def endpoint(self, worker: str) -> Endpoint:
        try:
            return self._endpoints[worker]
        except KeyError as exc:
            raise KeyError(f"No endpoint known for worker {worker!r}") from exc

    async def send(self, sender: str, recipient: str, payload: Any) -> None:
        await self.channel(sender).send(self.endpoint(recipient), payload)

    async def recv(self, worker: str, *, timeout: Optional[float] = None) -> Any:
        return await self.channel(worker).recv(timeout=timeout)


class ActorPool:
    """A set of worker actors that run :class:`~byzpy_b200.engine.graph.subtask.SubTask` objects.

    Parameters
    ----------
    configs : sequence of ActorPoolConfig
        Worker groups; a pool may mix backends (threads for cheap tasks, CUDA-stream workers for kernels, remote
        workers on other machines) and routes by ``SubTask.affinity``.

    Notes
    -----
    ``await pool.start()`` brings the workers up concurrently, ``await pool.shutdown()`` closes them; ``size`` is the
    number of workers (configured count before ``start``).  ``run_subtask`` / ``run_many`` hand each task to whichever idle
    worker has the capability it asks for; a task whose worker fails is retried up to ``max_retries`` times.
    ``open_channel(name)`` gives the workers named mailboxes to talk to each other.
    ``in_process`` is true when all workers share this address space, in which case operators hand views instead of
    shared-memory copies.  The pool pickles as its configuration only.

    Examples
    --------
    >>> import asyncio
    >>> from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig
    >>> from byzpy_b200.engine.graph.subtask import SubTask
    >>> async def demo():
    ...     pool = ActorPool([ActorPoolConfig("thread", count=2)])
    ...     await pool.start()
    ...     try:
    ...         return await pool.run_many([SubTask(fn=pow, args=(2, k)) for k in range(4)])
    ...     finally:
    ...         await pool.shutdown()
    >>> asyncio.run(demo())
    [1, 2, 4, 8]
    """

    def __init__(self, configs: Sequence[ActorPoolConfig]) -> None:
        self.configs = list(configs)
        self._workers: List[_PoolWorker] = []
        self._available: "asyncio.Queue[_PoolWorker]" = asyncio.Queue()
        self._waiting: Dict[Optional[str], Deque[asyncio.Future]] = defaultdict(deque)
        self._started = False
        self._channel_cache: Dict[str, ActorPoolChannel] = {}
        self._worker_affinity_caps: List[str] = []

    # a pool travels between processes as its configuration only (workers are re-created there)
    def __getstate__(self):
        return {"configs": self.configs}

    def __setstate__(self, state):
        self.__init__(state["configs"])

    @property
    def size(self) -> int:
        return len(self._workers) if self._started else sum(c.count for c in self.configs)

    @property
    def in_process(self) -> bool:
        """True when every worker shares this process's address space (thread / gpu backends):
        operators may then hand subtasks views of their inputs instead of shared-memory copies."""
        return all(isinstance(c.backend, str) and c.backend.split(":")[0] in ("thread", "gpu")
                   for c in self.configs)

    def worker_affinities(self) -> Sequence[str]:
        """The ``worker::<label>`` tag of every worker, for pinning subtasks to one of them."""
        return tuple(self._worker_affinity_caps)

    async def start(self) -> None:
        """Create and start the workers (idempotent); they come up concurrently."""
        if self._started:
            return
        fresh = []
        for cfg in self.configs:
            for idx in range(cfg.count):
                label = f"{cfg.name or 'actor'}-{idx}"
                pin = f"worker::{label}"
                worker = _PoolWorker(backend=resolve_backend(cfg.backend),
                                     capabilities=set(cfg.resolved_capabilities()) | {pin}, name=label)
                fresh.append((worker, pin))
        # workers come up concurrently (a process worker is a fresh interpreter importing torch: seconds each);
        # process workers split the host's cores between them (ProcessActorBackend sends the share with each call)
        await asyncio.gather(*(w.start() for w, _ in fresh))
        for worker, pin in fresh:
            self._workers.append(worker)
            self._worker_affinity_caps.append(pin)
            await self._release(worker)
        self._started = True

    async def shutdown(self) -> None:
        """Close every worker; the pool can be started again afterwards."""
        for w in self._workers:
            await w.close()
        self._workers.clear()
        self._started = False
        self._channel_cache.clear()
        self._worker_affinity_caps.clear()
        while not self._available.empty():
            self._available.get_nowait()
        for waiters in self._waiting.values():
            while waiters:
                fut = waiters.popleft()
                if not fut.done():
                    fut.set_exception(RuntimeError("ActorPool shutdown"))
        self._waiting.clear()

    async def open_channel(self, name: str) -> ActorPoolChannel:
        """A named mailbox on every worker: returns an :class:`ActorPoolChannel` through which workers (and the caller)
        exchange messages addressed by worker label.
        """
        await self.start()
        cached = self._channel_cache.get(name)
        if cached is not None:
            return cached
        chans = {w.name: await w.open_channel(name) for w in self._workers}
        eps = {w.name: await w.endpoint() for w in self._workers}
        self._channel_cache[name] = ActorPoolChannel(name=name, channels=chans, endpoints=eps)
        return self._channel_cache[name]

    async def run_many(self, subtasks: Sequence[SubTask]) -> List[Any]:
        """Run the subtasks concurrently (as far as workers are free) and return their results in the order given."""
        await self.start()
        if not subtasks:
            return []
        return list(await asyncio.gather(*[self._run_subtask(st) for st in subtasks]))

    async def run_subtask(self, subtask: SubTask) -> Any:
        """Run one subtask on an idle worker that has the capability ``subtask.affinity`` asks for; retries on another
        worker up to ``subtask.max_retries`` times when the worker raises.
        """
        await self.start()
        return await self._run_subtask(subtask)

    async def _run_subtask(self, subtask: SubTask) -> Any:
        attempts_left = max(0, subtask.max_retries) + 1
        while True:
            worker = await self._acquire(subtask.affinity)
            try:
                out = await worker.run(subtask)
                metrics.inc("byzpy_pool_subtasks_total")
                return out
            except Exception:
                attempts_left -= 1
                if attempts_left <= 0:
                    raise
                metrics.inc("byzpy_pool_subtask_retries_total")
            finally:
                await self._release(worker)

    async def _acquire(self, affinity: Optional[str]) -> _PoolWorker:
        if not self._workers:
            raise RuntimeError("ActorPool has no workers configured.")
        if affinity is None:
            if not self._available.empty():
                return self._available.get_nowait()
            fut = asyncio.get_running_loop().create_future()
            self._waiting[None].append(fut)
            return await self._await_worker(fut)
        # rotate through the idle workers once looking for a capable one
        for _ in range(self._available.qsize()):
            w = self._available.get_nowait()
            if affinity in w.capabilities:
                return w
            self._available.put_nowait(w)
        if not any(affinity in w.capabilities for w in self._workers):
            raise RuntimeError("No actor in the pool")
        fut = asyncio.get_running_loop().create_future()
        self._waiting[affinity].append(fut)
        return await self._await_worker(fut)

    async def _await_worker(self, fut: "asyncio.Future") -> _PoolWorker:
        """Wait for ``_release`` to hand over a worker.  A waiter that is cancelled in the same loop iteration in
        which it was handed one (the windowed runner cancels its in-flight subtasks when one of them fails) must
        give it back, or the pool shrinks by one worker for good."""
        try:
            return await fut
        except asyncio.CancelledError:
            if fut.done() and not fut.cancelled() and fut.exception() is None:
                await self._release(fut.result())
            raise

    async def _release(self, worker: _PoolWorker) -> None:
        for key in list(worker.capabilities) + [None]:
            waiters = self._waiting.get(key)
            while waiters:
                fut = waiters.popleft()
                if not fut.done():
                    fut.set_result(worker)
                    return
        self._available.put_nowait(worker)


__all__ = ["ActorPool", "ActorPoolConfig", "ActorPoolChannel"]
