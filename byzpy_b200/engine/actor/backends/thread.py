"""Thread actor: the object lives on a dedicated single-thread executor, so its methods never
run concurrently with each other (reference engine/actor/backends/thread.py:14-171)."""
from __future__ import annotations

import asyncio
import concurrent.futures
import inspect
from typing import Any

from ._local import LocalMailboxBackend, intra_op_governor


class ThreadActorBackend(LocalMailboxBackend):
    """Hosts the object on a dedicated single-thread executor of this process (``"thread"``).

    Method calls are serialised on that thread, so the object needs no locking of its own, and arguments and results are
    passed by reference (no pickling, CUDA tensors stay where they are).  When several thread actors compute at once
    the host's intra-op threads are divided between them (``BYZPY_INTRAOP_GOVERNOR=0`` disables that).

    Examples
    --------
    >>> import asyncio
    >>> from byzpy_b200.engine.actor.base import ActorRef
    >>> from byzpy_b200.engine.actor.backends.thread import ThreadActorBackend
    >>> async def demo():
    ...     be = ThreadActorBackend()
    ...     async with ActorRef(be) as ref:
    ...         await be.construct(list, args=([3, 1, 2],), kwargs={})
    ...         await ref.sort()
    ...         return await ref.copy()
    >>> asyncio.run(demo())
    [1, 2, 3]
    """

    scheme = "thread"

    def __init__(self) -> None:
        super().__init__()
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=1,
                                                           thread_name_prefix="byz-actor")
        self._obj: Any = None

    async def start(self) -> None:
        if self._loop is None:
            self._loop = asyncio.get_running_loop()

    async def _in_thread(self, fn, *args, **kwargs):
        loop = asyncio.get_running_loop()

        def governed():
            with intra_op_governor:
                return fn(*args, **kwargs)

        return await loop.run_in_executor(self._pool, governed)

    async def construct(self, cls_or_factory: Any, *, args: tuple, kwargs: dict) -> None:
        self._obj = await self._in_thread(cls_or_factory, *args, **kwargs)

    async def call(self, method: str, *args, **kwargs) -> Any:
        if self._obj is None:
            raise RuntimeError("actor not constructed")
        fn = getattr(self._obj, method)
        if inspect.iscoroutinefunction(fn):
            return await fn(*args, **kwargs)
        return await self._in_thread(fn, *args, **kwargs)

    async def close(self) -> None:
        self._unregister()
        self._pool.shutdown(wait=True)


__all__ = ["ThreadActorBackend"]
