"""Execution sessions with a node-name keyed result cache
(reference engine/graph/session.py:27-419).

``ExecutionSession.execute(graph, inputs)`` skips nodes whose *name* is already cached (cached
values are injected as graph inputs), always evaluates through :class:`ParallelScheduler`, and
records every newly computed node.  The cache is keyed by node name, not by input values, and is
cleared when the ``async with`` block exits.  ``execute_async`` returns an
:class:`ExecutionFuture`.
"""
from __future__ import annotations

import asyncio
from typing import Any, Dict, Mapping, Optional, Tuple

from .graph import ComputationGraph, GraphInput, GraphNode
from .parallel_scheduler import ParallelScheduler


class ExecutionFuture:
    """Handle on a graph execution started with :meth:`ExecutionSession.execute_async`.

    Awaitable (``await fut`` or ``await fut.wait()``) for the ``{output name: value}`` mapping; ``done()``,
    ``cancel()``, ``cancelled()`` as on an asyncio task; ``result(timeout=None)`` returns the mapping once finished
    (raises when it is not); ``output_keys`` names what will be returned.
    """

    def __init__(self, task: "asyncio.Task[Dict[str, Any]]", output_keys) -> None:
        self._task = task
        self._output_keys = tuple(output_keys)

    @property
    def output_keys(self) -> Tuple[str, ...]:
        return self._output_keys

    def done(self) -> bool:
        return self._task.done()

    def cancel(self) -> bool:
        return self._task.cancel()

    def cancelled(self) -> bool:
        return self._task.cancelled()

    async def wait(self) -> Dict[str, Any]:
        return await self._task

    def __await__(self):
        return self._task.__await__()

    def result(self, timeout: Optional[float] = None) -> Dict[str, Any]:
        try:
            loop = asyncio.get_running_loop()
        except RuntimeError:
            loop = None
        if loop is not None and loop.is_running():
            raise RuntimeError("Cannot call result() from within a running event loop. "
                               "Use 'await future' instead.")
        if self._task.done():
            return self._task.result()
        owner = self._task.get_loop()
        return owner.run_until_complete(asyncio.wait_for(self._task, timeout=timeout))


class ExecutionSession:
    """Runs graphs through a :class:`~byzpy_b200.engine.graph.parallel_scheduler.ParallelScheduler`, remembering results by node name.

    Parameters
    ----------
    pool : ActorPool, optional
    cache_intermediate : bool, default True
        Keep every computed node's result; a later ``execute`` skips nodes whose *name* is cached (the cache does not look
        at input values -- give nodes fed with different data different names, or call ``clear_cache``).
    metadata : mapping, optional

    Notes
    -----
    Use as ``async with ExecutionSession(pool) as s: ...``; the cache is dropped on exit.  ``get_cached`` /
    ``is_cached`` inspect it.

    Examples
    --------
    >>> import asyncio, torch
    >>> from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
    >>> from byzpy_b200.engine.graph.ops import make_single_operator_graph
    >>> from byzpy_b200.engine.graph.session import ExecutionSession
    >>> graph = make_single_operator_graph(node_name="agg", operator=CoordinateWiseMedian(), input_keys=("gradients",))
    >>> async def demo():
    ...     async with ExecutionSession() as s:
    ...         first = await s.execute(graph, {"gradients": [torch.tensor([1.0]), torch.tensor([3.0]), torch.tensor([2.0])]})
    ...         return first, s.is_cached("agg")
    >>> asyncio.run(demo())
    ({'agg': tensor([2.])}, True)
    """

    def __init__(self, pool=None, cache_intermediate: bool = True,
                 metadata: Optional[Mapping[str, Any]] = None) -> None:
        self.pool = pool
        self.cache_intermediate = cache_intermediate
        self.metadata = dict(metadata or {})
        self._result_cache: Dict[str, Any] = {}

    async def __aenter__(self) -> "ExecutionSession":
        return self

    async def __aexit__(self, exc_type, exc_val, exc_tb) -> None:
        self._result_cache.clear()

    async def execute(self, graph: ComputationGraph, inputs: Mapping[str, Any]) -> Dict[str, Any]:
        """Run ``graph`` on ``inputs``; nodes whose names are cached are not recomputed.  Returns ``{output: value}``."""
        if not self.cache_intermediate:
            return await ParallelScheduler(graph, pool=self.pool, metadata=self.metadata).run(inputs)
        pruned, cached = self._prune_cached_nodes(graph)
        if pruned is None:
            return {name: self._result_cache[name] for name in graph.outputs}
        todo = list(pruned.nodes_in_order())
        everything = ComputationGraph(todo, outputs=[n.name for n in todo])
        feed = dict(inputs)
        feed.update(cached)
        fresh = await ParallelScheduler(everything, pool=self.pool, metadata=self.metadata).run(feed)
        self._result_cache.update(fresh)
        return {name: self._result_cache[name] for name in graph.outputs}

    def execute_async(self, graph: ComputationGraph, inputs: Mapping[str, Any]) -> ExecutionFuture:
        """Start :meth:`execute` in the background and return an :class:`ExecutionFuture`."""
        task = asyncio.ensure_future(self.execute(graph, inputs))
        return ExecutionFuture(task, output_keys=graph.outputs)

    def _prune_cached_nodes(self, graph: ComputationGraph):
        keep = []
        cached: Dict[str, Any] = {}
        for node in graph.nodes_in_order():
            if node.name in self._result_cache:
                cached[node.name] = self._result_cache[node.name]
                continue
            rewired = {}
            for arg, dep in node.inputs.items():
                if isinstance(dep, str) and dep in self._result_cache:
                    rewired[arg] = GraphInput(dep)
                else:
                    rewired[arg] = dep
            keep.append(GraphNode(name=node.name, op=node.op, inputs=rewired))
        if not keep:
            return None, cached
        outs = [name for name in graph.outputs if name not in self._result_cache]
        if not outs:
            return None, cached
        return ComputationGraph(keep, outputs=outs), cached

    def clear_cache(self) -> None:
        """Forget every cached node result."""
        self._result_cache.clear()

    def get_cached(self, key: str) -> Any:
        """Cached result of node ``key`` (``KeyError`` when absent)."""
        if key not in self._result_cache:
            raise KeyError(f"No cached result for {key!r}")
        return self._result_cache[key]

    def is_cached(self, key: str) -> bool:
        """Whether node ``key`` has a cached result."""
        return key in self._result_cache


__all__ = ["ExecutionSession", "ExecutionFuture"]
