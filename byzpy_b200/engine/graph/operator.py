"""Operator base classes of the scheduling layer.

An ``Operator`` is a node payload of a :class:`ComputationGraph`.  It can run
(a) directly (``compute``), (b) decomposed into independent ``SubTask`` s that a
pool executes with a sliding in-flight window, or (c) as a *barriered* iterative
procedure that drives the pool itself.  The dispatch contract matches the
reference (reference engine/graph/operator.py:55-70):

    barriered  if supports_barriered_subtasks and a pool is given
    subtasks   elif supports_subtasks and pool.size > 1   (falls through to compute
               when no partials were produced)
    compute    otherwise

All operators in this framework are stateless across invocations (workspaces are
passed explicitly), so one instance may run concurrently in several graph nodes --
the reference keeps per-call state on ``self`` and is not re-entrant.
"""
from __future__ import annotations

import asyncio
import dataclasses
import inspect
from dataclasses import dataclass
from typing import TYPE_CHECKING, Any, Iterable, Iterator, List, Mapping, Optional, Sequence

from .subtask import SubTask

if TYPE_CHECKING:  # pragma: no cover
    from .pool import ActorPool


@dataclass(frozen=True)
class OpContext:
    """Runtime metadata handed to each operator invocation."""

    node_name: str
    metadata: Mapping[str, Any] | None = None


async def _resolve(value: Any) -> Any:
    return (await value) if inspect.isawaitable(value) else value


def _window_size(limit: Optional[int], pool_size: int) -> int:
    if not limit:  # None or 0 -> default window
        return max(1, pool_size * 8)
    if limit < 0:
        return max(1, pool_size * (-limit))
    return limit


def _with_affinities(subtasks: Iterable[SubTask], hints: Sequence[str]) -> Iterator[SubTask]:
    k = 0
    for st in subtasks:
        if st.affinity is None and hints:
            st = dataclasses.replace(st, affinity=hints[k % len(hints)])
        k += 1
        yield st


async def run_subtasks_windowed(pool: "ActorPool", subtasks: Iterable[SubTask],
                                limit: Optional[int],
                                semaphore: Optional[asyncio.Semaphore] = None) -> List[Any]:
    """Run ``subtasks`` keeping at most ``window`` in flight; results keep submission order.

    ``semaphore`` (optional, shared across operators of one scheduler) additionally
    bounds the total number of pending subtasks.
    """
    window = _window_size(limit, pool.size)
    source = iter(subtasks)
    inflight: set = set()
    results: dict = {}
    submitted = 0

    async def _one(st: SubTask, slot: int):
        try:
            return slot, await pool.run_subtask(st)
        finally:
            if semaphore is not None:
                semaphore.release()

    async def _submit() -> bool:
        nonlocal submitted
        st = next(source, None)
        if st is None:
            return False
        if semaphore is not None:
            await semaphore.acquire()
        try:
            task = asyncio.ensure_future(_one(st, submitted))
        except BaseException:
            if semaphore is not None:
                semaphore.release()
            raise
        inflight.add(task)
        submitted += 1
        return True

    while len(inflight) < window and await _submit():
        pass
    try:
        while inflight:
            done, inflight = await asyncio.wait(inflight, return_when=asyncio.FIRST_COMPLETED)
            for t in done:
                slot, value = t.result()
                results[slot] = value
            while len(inflight) < window and await _submit():
                pass
    except BaseException:
        for t in inflight:
            t.cancel()
        raise
    return [results[i] for i in range(submitted)]


class Operator:
    """Base class of everything that can sit in a computation graph."""

    name: str = "operator"
    supports_subtasks: bool = False
    supports_barriered_subtasks: bool = False
    max_subtasks_inflight: Optional[int] = None
    # The reference's operators park per-call state on ``self`` between create_subtasks() and
    # reduce_subtasks() (a shared-memory handle and the flat shape).  Operators here are stateless, the
    # attributes exist (always None) so callers that inspect or reset them keep working.
    _active_handle: Any = None
    _flat_shape: Any = None

    # -- to be provided by subclasses -------------------------------------------------
    def compute(self, inputs: Mapping[str, Any], *, context: OpContext) -> Any:
        raise NotImplementedError

    def create_subtasks(self, inputs: Mapping[str, Any], *, context: OpContext) -> Iterable[SubTask]:
        return []

    def reduce_subtasks(self, partials: Sequence[Any], inputs: Mapping[str, Any], *,
                        context: OpContext) -> Any:
        raise RuntimeError(f"Operator {self.name} does not implement reduce_subtasks().")

    async def run_barriered_subtasks(self, inputs: Mapping[str, Any], *, context: OpContext,
                                     pool: "ActorPool") -> Any:
        raise RuntimeError(f"Operator {self.name} does not implement barriered subtasks.")

    # -- dispatch ------------------------------------------------------------------------
    async def run(self, inputs: Mapping[str, Any], *, context: OpContext,
                  pool: Optional["ActorPool"]) -> Any:
        if pool is not None and self.supports_barriered_subtasks:
            return await _resolve(self.run_barriered_subtasks(inputs, context=context, pool=pool))
        if pool is not None and self.supports_subtasks and pool.size > 1:
            partials = await self._run_subtasks(
                pool, self.create_subtasks(inputs, context=context), self.max_subtasks_inflight,
                context)
            if partials:
                return await _resolve(self.reduce_subtasks(partials, inputs, context=context))
        if (context.metadata or {}).get("offload_host_compute"):
            # ParallelScheduler, several nodes in flight: a synchronous compute() would hold the event
            # loop and serialise the wave, so it runs on a worker thread (torch / the native kernels
            # release the GIL), sharing the cores with its siblings through the intra-op governor
            return await _resolve(await asyncio.to_thread(_governed_compute, self, inputs, context))
        return await _resolve(self.compute(inputs, context=context))

    async def _run_subtasks(self, pool: "ActorPool", subtasks: Iterable[SubTask],
                            limit: Optional[int], context: OpContext) -> List[Any]:
        meta = context.metadata or {}
        hints = meta.get("worker_affinities")
        if hints:
            subtasks = _with_affinities(subtasks, tuple(hints))
        return await run_subtasks_windowed(pool, subtasks, limit, meta.get("subtask_semaphore"))


def _governed_compute(op: "Operator", inputs: Mapping[str, Any], context: OpContext) -> Any:
    from ..actor.backends._local import intra_op_governor

    with intra_op_governor:
        return op.compute(inputs, context=context)


class MessageTriggerOp(Operator):
    """Blocks until the owning scheduler receives a message of ``message_type``."""

    def __init__(self, message_type: str, timeout: Optional[float] = None):
        if not message_type:
            raise ValueError("message_type cannot be empty")
        self.message_type = message_type
        self.timeout = timeout
        self.name = f"message_trigger_{message_type}"

    async def run(self, inputs: Mapping[str, Any], *, context: OpContext,
                  pool: Optional["ActorPool"]) -> Any:
        scheduler = (context.metadata or {}).get("scheduler")
        if scheduler is None:
            raise RuntimeError("MessageTriggerOp requires scheduler in context metadata")
        return await scheduler.wait_for_message(self.message_type, timeout=self.timeout)


# private aliases kept for parity with the reference's helper names
_run_subtasks_windowed = run_subtasks_windowed
_maybe_await = _resolve

__all__ = ["OpContext", "Operator", "MessageTriggerOp", "run_subtasks_windowed"]
