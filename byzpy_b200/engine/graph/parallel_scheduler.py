"""Dataflow scheduler: independent graph nodes run concurrently.

API of the reference scheduler (reference engine/graph/parallel_scheduler.py:19-278):
``ParallelScheduler(graph, pool=None, metadata=None, max_concurrent_nodes=None,
max_pending_subtasks=None)``; a shared subtask semaphore (default ``pool.size * 8``, ``0``
disables it) bounds the pending subtasks of all concurrently running operators.

Differences by design:
* true dataflow -- a node starts the moment its own dependencies finish (the reference runs in
  waves: a whole wave must finish before any dependent starts);
* **CUDA-stream dispatch** -- when a node's inputs are CUDA tensors, the node body is issued on
  its own CUDA stream taken from a small pool, ordered after its producers by CUDA events
  instead of host-side waits, so independent branches overlap ON THE DEVICE (graph nodes map to
  streams rather than to thread/process actors).  Disable with ``metadata={"cuda_streams": False}``.
  The side stream is made current around the node's ``run`` call, and "current stream" is a property of
  the host THREAD: an operator whose ``compute`` is a coroutine that suspends (awaits I/O, another
  actor, a message) hands the thread -- with the side stream still current -- to whichever node the
  event loop resumes next.  Stream dispatch is therefore meant for operators that enqueue their device
  work synchronously (every operator of this library does); give asynchronous operators host inputs or
  switch the streams off for that graph.
"""
from __future__ import annotations

import asyncio
from collections import defaultdict
from typing import Any, Dict, List, Mapping, Optional, Tuple

from .graph import ComputationGraph, GraphInput, GraphNode
from .operator import OpContext


def _cuda_tensors(obj: Any, out: List[Any], depth: int = 0) -> None:
    try:
        import torch
    except Exception:  # pragma: no cover
        return
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            out.append(obj)
    elif depth < 3 and isinstance(obj, (list, tuple)):
        for x in obj[:256]:
            _cuda_tensors(x, out, depth + 1)
    elif depth < 3 and isinstance(obj, dict):
        for x in obj.values():
            _cuda_tensors(x, out, depth + 1)


class _StreamPool:
    """Round-robin pool of side streams per device."""

    def __init__(self, size: int = 4) -> None:
        self.size = size
        self._streams: Dict[int, list] = {}
        self._next: Dict[int, int] = defaultdict(int)

    def take(self, device):
        import torch

        idx = device.index if device.index is not None else torch.cuda.current_device()
        pool = self._streams.get(idx)
        if pool is None:
            pool = [torch.cuda.Stream(device=idx) for _ in range(self.size)]
            self._streams[idx] = pool
        s = pool[self._next[idx] % self.size]
        self._next[idx] += 1
        return s


class ParallelScheduler:
    """Dataflow scheduler: every node starts as soon as its inputs are ready, independent branches run concurrently.

    Parameters
    ----------
    graph : ComputationGraph
    pool : ActorPool, optional
        Shared by all concurrently running operators.
    metadata : mapping, optional
        As for :class:`~byzpy_b200.engine.graph.scheduler.NodeScheduler`.
    max_concurrent_nodes : int, optional
        Upper bound on nodes in flight.
    max_pending_subtasks : int, optional
        Upper bound on subtasks in flight over all running operators; default ``8 x pool.size``.

    Notes
    -----
    Nodes whose inputs are CUDA tensors are issued on CUDA streams taken from a per-device stream pool, ordered by
    events instead of host synchronisation, so independent branches overlap on the GPU as well as on the host.

    Examples
    --------
    >>> import asyncio, torch
    >>> from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean
    >>> from byzpy_b200.engine.graph.lazy import GraphBuilder
    >>> from byzpy_b200.engine.graph.parallel_scheduler import ParallelScheduler
    >>> b = GraphBuilder()
    >>> x = b.input("gradients")
    >>> med = x.apply(CoordinateWiseMedian(), name="median")
    >>> tm = x.apply(CoordinateWiseTrimmedMean(f=1), name="trimmed")
    >>> grads = [torch.tensor([v]) for v in (1.0, 2.0, 6.0, 100.0)]
    >>> out = asyncio.run(ParallelScheduler(b.build(outputs=["median", "trimmed"])).run({"gradients": grads}))
    >>> out["median"], out["trimmed"]
    (tensor([2.]), tensor([4.]))
    """

    def __init__(self, graph: ComputationGraph, *, pool=None,
                 metadata: Optional[Mapping[str, Any]] = None,
                 max_concurrent_nodes: Optional[int] = None,
                 max_pending_subtasks: Optional[int] = None) -> None:
        self.graph = graph
        self.pool = pool
        self.metadata = dict(metadata or {})
        self.max_concurrent_nodes = max_concurrent_nodes
        if max_pending_subtasks is None and pool is not None:
            max_pending_subtasks = pool.size * 8
        self.max_pending_subtasks = max_pending_subtasks
        self._node_map: Dict[str, GraphNode] = {n.name: n for n in graph.nodes_in_order()}
        self._in_degree: Dict[str, int] = {}
        self._dependents: Dict[str, List[str]] = defaultdict(list)
        for name in self._node_map:
            deps = graph.dependencies(name)
            self._in_degree[name] = len(deps)
            for dep in deps:
                self._dependents[dep].append(name)
        self._streams = _StreamPool()

    # ------------------------------------------------------------------------------
    def _resolve_inputs(self, node: GraphNode, cache: Dict[str, Any]) -> Dict[str, Any]:
        bound: Dict[str, Any] = {}
        for arg, dep in node.inputs.items():
            if isinstance(dep, GraphInput):
                bound[arg] = cache[dep.name]
            elif isinstance(dep, str):
                if dep not in cache:
                    raise KeyError(f"Graph node {node.name} depends on {dep!r}, which has not been computed.")
                bound[arg] = cache[dep]
            else:
                bound[arg] = dep
        return bound

    async def _execute_node(self, node_name: str, cache: Dict[str, Any], base_metadata: Dict[str, Any],
                            events: Optional[Dict[str, Any]] = None, concurrent: bool = False) -> Tuple[str, Any]:
        node = self._node_map[node_name]
        inputs = self._resolve_inputs(node, cache)
        device_inputs: List[Any] = []
        _cuda_tensors(inputs, device_inputs)
        # worker-thread offload is for host data only: the current CUDA stream is thread-local, so device
        # work stays on this thread (and goes onto a side stream below when no pool is attached)
        offload = concurrent and not device_inputs and bool(base_metadata.get("offload_host_compute", True))
        if offload != bool(base_metadata.get("offload_host_compute")):
            base_metadata = {**base_metadata, "offload_host_compute": offload}
        ctx = OpContext(node_name=node.name, metadata=base_metadata)
        use_streams = events is not None and base_metadata.get("cuda_streams", True) and self.pool is None
        tensors: List[Any] = device_inputs if use_streams else []
        tracer = base_metadata.get("tracer")
        if not tensors:
            if tracer is None:
                return node_name, await node.op.run(inputs, context=ctx, pool=self.pool)
            with tracer.span(f"node:{node_name}", op=node.op.name):
                return node_name, await node.op.run(inputs, context=ctx, pool=self.pool)
        import torch

        dev = tensors[0].device
        stream = self._streams.take(dev)
        # order after the launching stream and after every producer node (device-side waits only)
        stream.wait_stream(torch.cuda.current_stream(dev))
        for dep in self.graph.dependencies(node_name):
            ev = events.get(dep)
            if ev is not None:
                stream.wait_event(ev)
        for t in tensors:
            t.record_stream(stream)
        with torch.cuda.stream(stream):
            result = await node.op.run(inputs, context=ctx, pool=self.pool)
            done = torch.cuda.Event()
            done.record(stream)
        events[node_name] = done
        events.setdefault("__streams__", []).append(stream)
        return node_name, result

    async def run(self, inputs: Mapping[str, Any]) -> Dict[str, Any]:
        """Execute the graph on ``inputs`` with every ready node in flight at once; returns ``{output node name: value}``.
        A failing node cancels the rest and its exception propagates.
        """
        missing = [name for name in self.graph.required_inputs if name not in inputs]
        if missing:
            raise ValueError(f"Missing graph inputs: {missing}")
        cache: Dict[str, Any] = dict(inputs)
        remaining = dict(self._in_degree)
        meta = dict(self.metadata)
        if self.pool is not None:
            meta.setdefault("pool_size", self.pool.size)
            meta.setdefault("pool_in_process", bool(getattr(self.pool, "in_process", False)))
            meta.setdefault("worker_affinities", tuple(self.pool.worker_affinities()))
        if self.max_pending_subtasks:
            meta["subtask_semaphore"] = asyncio.Semaphore(self.max_pending_subtasks)
        gate = (asyncio.Semaphore(self.max_concurrent_nodes)
                if self.max_concurrent_nodes and self.max_concurrent_nodes > 0 else None)
        events: Dict[str, Any] = {}

        async def guarded(name: str, concurrent: bool = False):
            if gate is None:
                return await self._execute_node(name, cache, meta, events, concurrent)
            async with gate:
                return await self._execute_node(name, cache, meta, events, concurrent)

        running: set = set()
        ready = [name for name in self._node_map if remaining[name] == 0]
        try:
            while ready or running:
                if len(ready) == 1 and not running:
                    # fast path: a lone ready node runs inline (no task overhead)
                    name, value = await guarded(ready.pop())
                    finished = [(name, value)]
                else:
                    # more than one node in flight: host-side compute() calls go to worker threads
                    together = len(ready) + len(running) > 1
                    for name in ready:
                        running.add(asyncio.ensure_future(guarded(name, together)))
                    ready = []
                    done, running = await asyncio.wait(running, return_when=asyncio.FIRST_COMPLETED)
                    finished = [t.result() for t in done]
                for name, value in finished:
                    cache[name] = value
                    for child in self._dependents[name]:
                        remaining[child] -= 1
                        if remaining[child] == 0:
                            ready.append(child)
        except BaseException:
            for t in running:
                t.cancel()
            raise
        # join the side streams back into the caller's stream (device-side)
        side = events.get("__streams__")
        if side:
            import torch

            for s in set(side):
                torch.cuda.current_stream(s.device).wait_stream(s)
        return {name: cache[name] for name in self.graph.outputs}


__all__ = ["ParallelScheduler"]
