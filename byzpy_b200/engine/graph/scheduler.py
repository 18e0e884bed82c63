"""Sequential graph schedulers.

``NodeScheduler`` evaluates a :class:`ComputationGraph` in topological order, handing each
operator the optional :class:`ActorPool` (reference engine/graph/scheduler.py:12-81).
``MessageAwareNodeScheduler`` adds per-type message queues so graph inputs / node inputs can be
``MessageSource`` s and ``MessageTriggerOp`` can block on deliveries (reference
scheduler.py:84-269).  ``deliver_message`` wakes every current waiter AND queues the payload
(reference scheduler.py:188-194).

On CUDA inputs the per-node work is just asynchronous kernel launches on the current stream, so
a sequential scheduler already keeps the GPU queue full; inter-node concurrency on separate
CUDA streams is provided by :class:`ParallelScheduler`.
"""
from __future__ import annotations

import asyncio
from collections import defaultdict
from typing import Any, Dict, List, Mapping, MutableMapping, Optional

from .graph import ComputationGraph, GraphInput, GraphNode
from .operator import OpContext


class MessageSource:
    """Graph/node input that is filled from the next message of ``message_type``."""

    def __init__(self, message_type: str, field: Optional[str] = None,
                 timeout: Optional[float] = None):
        self.message_type = message_type
        self.field = field
        self.timeout = timeout

    def __repr__(self) -> str:
        return f"MessageSource({self.message_type!r}, field={self.field!r}, timeout={self.timeout!r})"


class NodeScheduler:
    """Runs a :class:`~byzpy_b200.engine.graph.graph.ComputationGraph` node by node in topological order.

    Parameters
    ----------
    graph : ComputationGraph
    pool : ActorPool, optional
        Given a pool, operators that support subtasks may fan their work out to it (``Operator.run`` decides, see
        ``BYZPY_POOL_DISPATCH``); without one every operator computes in the calling task.
    metadata : mapping, optional
        Merged into every operator's ``OpContext.metadata`` (``pool_size``, ``pool_in_process`` and
        ``worker_affinities`` are filled in from the pool).

    Examples
    --------
    >>> import asyncio, torch
    >>> from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
    >>> from byzpy_b200.engine.graph.ops import make_single_operator_graph
    >>> from byzpy_b200.engine.graph.scheduler import NodeScheduler
    >>> graph = make_single_operator_graph(node_name="agg", operator=CoordinateWiseMedian(), input_keys=("gradients",))
    >>> grads = [torch.tensor([1.0]), torch.tensor([5.0]), torch.tensor([2.0])]
    >>> asyncio.run(NodeScheduler(graph).run({"gradients": grads}))
    {'agg': tensor([2.])}
    """

    def __init__(self, graph: ComputationGraph, *, pool=None,
                 metadata: Optional[Mapping[str, Any]] = None) -> None:
        self.graph = graph
        self.pool = pool
        self.metadata = dict(metadata or {})

    def _check_inputs(self, inputs: Mapping[str, Any]) -> None:
        missing = [name for name in self.graph.required_inputs if name not in inputs]
        if missing:
            raise ValueError(f"Missing graph inputs: {missing}")

    def _node_metadata(self) -> Dict[str, Any]:
        meta = dict(self.metadata)
        if self.pool is not None:
            meta.setdefault("pool_size", self.pool.size)
            meta.setdefault("pool_in_process", bool(getattr(self.pool, "in_process", False)))
            meta.setdefault("worker_affinities", tuple(self.pool.worker_affinities()))
        return meta

    def _resolve_inputs(self, node: GraphNode, cache: MutableMapping[str, Any]) -> Dict[str, Any]:
        bound: Dict[str, Any] = {}
        for arg, dep in node.inputs.items():
            if isinstance(dep, GraphInput):
                bound[arg] = cache[dep.name]
            elif dep in cache:
                bound[arg] = cache[dep]
            else:
                raise KeyError(f"Graph node {node.name} depends on {dep!r}, which has not been computed.")
        return bound

    async def run(self, inputs: Mapping[str, Any]) -> Dict[str, Any]:
        """Execute the graph on ``inputs`` (a mapping covering ``graph.required_inputs``) and return
        ``{output node name: value}``.  Raises ``ValueError`` when an input is missing.
        """
        self._check_inputs(inputs)
        cache: Dict[str, Any] = dict(inputs)
        tracer = self.metadata.get("tracer")
        for node in self.graph.nodes_in_order():
            ctx = OpContext(node_name=node.name, metadata=self._node_metadata())
            bound = self._resolve_inputs(node, cache)
            if tracer is None:
                cache[node.name] = await node.op.run(bound, context=ctx, pool=self.pool)
            else:
                with tracer.span(f"node:{node.name}", op=node.op.name):
                    cache[node.name] = await node.op.run(bound, context=ctx, pool=self.pool)
        return {name: cache[name] for name in self.graph.outputs}


class MessageAwareNodeScheduler(NodeScheduler):
    """A :class:`NodeScheduler` whose graphs may take inputs from messages that arrive while the graph runs.

    An input declared as ``GraphInput.from_message("gradient", field="vector", timeout=5)`` makes the node wait until
    somebody calls ``deliver_message("gradient", payload)`` (a decentralized node's message loop does); messages that
    arrive before anyone waits are queued per type.  ``wait_for_message`` is the same primitive for hand-written
    pipelines.

    Examples
    --------
    >>> import asyncio
    >>> from byzpy_b200.engine.graph.graph import ComputationGraph, GraphInput, GraphNode
    >>> from byzpy_b200.engine.graph.ops import CallableOp
    >>> from byzpy_b200.engine.graph.scheduler import MessageAwareNodeScheduler
    >>> g = ComputationGraph([GraphNode("double", CallableOp(lambda x: 2 * x, input_mapping={"x": "x"}),
    ...                                 {"x": GraphInput.from_message("number", field="value")})])
    >>> async def demo():
    ...     sched = MessageAwareNodeScheduler(g)
    ...     task = asyncio.ensure_future(sched.run({}))
    ...     await asyncio.sleep(0)
    ...     sched.deliver_message("number", {"value": 21})
    ...     return await task
    >>> asyncio.run(demo())
    {'double': 42}
    """

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._message_waiters: Dict[str, List[asyncio.Future]] = {}
        self._message_cache: Dict[str, List[Any]] = defaultdict(list)

    async def wait_for_message(self, message_type: str, *, timeout: Optional[float] = None) -> Any:
        """Next message of ``message_type``: a queued one if any, else wait up to ``timeout`` seconds
        (``asyncio.TimeoutError``).
        """
        queued = self._message_cache.get(message_type)
        if queued:
            return queued.pop(0)
        fut = asyncio.get_running_loop().create_future()
        self._message_waiters.setdefault(message_type, []).append(fut)
        try:
            return await asyncio.wait_for(fut, timeout=timeout)
        except asyncio.TimeoutError:
            pending = self._message_waiters.get(message_type, [])
            if fut in pending:
                pending.remove(fut)
            raise

    def deliver_message(self, message_type: str, payload: Any) -> None:
        """Hand a message to the scheduler: wakes every task waiting for ``message_type`` and queues the payload for later
        waiters.
        """
        for fut in self._message_waiters.pop(message_type, []):
            if not fut.done():
                fut.set_result(payload)
        self._message_cache[message_type].append(payload)

    async def _from_message(self, src: MessageSource) -> Any:
        msg = await self.wait_for_message(src.message_type, timeout=src.timeout)
        if not src.field:
            return msg
        if not isinstance(msg, dict):
            raise TypeError(f"Cannot extract field '{src.field}' from non-dict message")
        if src.field not in msg:
            raise KeyError(f"Message field '{src.field}' not found in message payload")
        return msg[src.field]

    async def _resolve_inputs(self, node: GraphNode, cache: MutableMapping[str, Any]) -> Dict[str, Any]:  # type: ignore[override]
        bound: Dict[str, Any] = {}
        for arg, dep in node.inputs.items():
            if isinstance(dep, MessageSource):
                bound[arg] = await self._from_message(dep)
            elif isinstance(dep, GraphInput):
                bound[arg] = cache[dep.name]
            elif dep in cache:
                bound[arg] = cache[dep]
            else:
                raise KeyError(f"Graph node {node.name} depends on {dep!r}, which has not been computed.")
        return bound

    async def run(self, inputs: Mapping[str, Any]) -> Dict[str, Any]:
        resolved: Dict[str, Any] = {}
        for key, value in inputs.items():
            resolved[key] = await self._from_message(value) if isinstance(value, MessageSource) else value
        self._check_inputs(resolved)
        cache: Dict[str, Any] = dict(resolved)
        for node in self.graph.nodes_in_order():
            meta = self._node_metadata()
            meta["scheduler"] = self
            ctx = OpContext(node_name=node.name, metadata=meta)
            cache[node.name] = await node.op.run(await self._resolve_inputs(node, cache), context=ctx,
                                                 pool=self.pool)
        return {name: cache[name] for name in self.graph.outputs}


__all__ = ["NodeScheduler", "MessageAwareNodeScheduler", "MessageSource"]
