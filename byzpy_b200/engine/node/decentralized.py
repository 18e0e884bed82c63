"""``DecentralizedNode``: scheduler + application + message routing for one autonomous node
(reference engine/node/decentralized.py:12-284).

Incoming messages are (1) delivered to the node's :class:`MessageAwareNodeScheduler` so that
pipelines waiting on a ``MessageSource`` / ``MessageTriggerOp`` wake up, and (2) dispatched to
the handler registered for the message type.  ``execute_pipeline`` runs one of the application's
pipelines on the node's scheduler; ``start_autonomous_task`` parks named background coroutines
that are cancelled on shutdown.
"""
from __future__ import annotations

import asyncio
from typing import Any, Awaitable, Callable, Dict, List, Mapping, Optional, Union

from ...utils import metrics
from ..graph.graph import ComputationGraph, GraphNode
from ..graph.operator import Operator
from ..graph.scheduler import MessageAwareNodeScheduler
from .application import NodeApplication
from .context import NodeContext
from .router import MessageRouter

NodeId = Union[int, str]


class _Idle(Operator):
    name = "idle"

    def compute(self, inputs, *, context):
        return None


async def _cancel(task: Optional[asyncio.Task], timeout: float = 1.0) -> None:
    if task is None or task.done():
        return
    task.cancel()
    try:
        await asyncio.wait_for(task, timeout=timeout)
    except (asyncio.CancelledError, asyncio.TimeoutError, Exception):
        pass


class DecentralizedNode:
    """One autonomous participant: an application (pool + pipelines), a message-aware scheduler, a router and a context.

    Parameters
    ----------
    node_id : int or str
    application : NodeApplication
        What the node can compute.
    context : NodeContext
        Where it lives and how its messages travel.
    topology : Topology, optional
        Restricts whom it may talk to (``None``: anybody, and broadcasts reach nobody).
    metadata : mapping, optional
    node_id_map : dict, optional
        Topology index -> node id, when ids are strings.

    Notes
    -----
    ``await start()`` / ``await shutdown()``.  Outbound: ``send_message``, ``broadcast_message`` (out-neighbours),
    ``multicast_message``.  Inbound messages wake pipelines waiting on that message type and are passed to the handler
    registered with ``register_message_handler(type, async_fn(sender, payload))``; a handler that raises is recorded
    in ``handler_errors`` and does not stop message processing.  ``await execute_pipeline(name, inputs)`` runs one of
    the application's pipelines on the node's scheduler; ``await start_autonomous_task(coro, name)`` parks a background
    task (a training loop) that is cancelled on shutdown.  See :class:`~byzpy_b200.engine.node.context.InProcessContext`
    for a runnable example.
    """

    def __init__(self, *, node_id: NodeId, application: NodeApplication, context: NodeContext,
                 topology: Optional[Any] = None, metadata: Optional[Mapping[str, Any]] = None,
                 node_id_map: Optional[Dict[int, str]] = None):
        if node_id is None or (isinstance(node_id, str) and node_id == ""):
            raise ValueError("node_id cannot be empty")
        self.node_id = node_id
        self.application = application
        self.context = context
        self.topology = topology
        self._node_id_map = node_id_map
        meta = dict(metadata or {})
        meta["node_id"] = node_id
        placeholder = ComputationGraph([GraphNode(name="dummy", op=_Idle(), inputs={})], outputs=["dummy"])
        self.scheduler = MessageAwareNodeScheduler(graph=placeholder, pool=application.pool, metadata=meta)
        self.message_router = MessageRouter(topology=topology, node_id=node_id, node_id_map=node_id_map)
        self._state: Dict[str, Any] = {}
        self._message_handlers: Dict[str, Callable[[str, Any], Awaitable[None]]] = {}
        self._running = False
        self._message_task: Optional[asyncio.Task] = None
        self._autonomous_tasks: Dict[str, asyncio.Task] = {}
        self.handler_errors: List[tuple] = []          # (message type, repr(exception)) of handlers that raised
        self._register_default_handlers()

    # ---- lifecycle ------------------------------------------------------------------------
    async def start(self) -> None:
        """Attach the context and start processing incoming messages (idempotent)."""
        if self._running:
            return
        self._running = True
        await self.context.start(self)
        self._message_task = asyncio.ensure_future(self._message_processing_loop())

    async def shutdown(self) -> None:
        """Cancel autonomous tasks and message processing, shut the context and the application's pool down."""
        if not self._running:
            return
        self._running = False
        for task in list(self._autonomous_tasks.values()):
            await _cancel(task)
        self._autonomous_tasks.clear()
        await _cancel(self._message_task)
        await self.context.shutdown()
        await self.application.shutdown()

    # ---- inbound ----------------------------------------------------------------------------
    async def _message_processing_loop(self) -> None:
        try:
            async for msg in self.context.receive_messages():
                if not self._running:
                    break
                try:
                    await self.handle_incoming_message(from_node_id=msg.get("from", "unknown"),
                                                       message_type=msg.get("type", "unknown"),
                                                       payload=msg.get("payload"))
                except asyncio.CancelledError:
                    raise
                except Exception as exc:  # noqa: BLE001
                    # a raising handler must not end message processing for the node (in the reference the loop
                    # task dies with the exception and the node goes deaf without a trace, reference
                    # engine/node/decentralized.py:95-107): record it, say so once per message type, carry on
                    mtype = msg.get("type", "unknown")
                    first = not any(t == mtype for t, _ in self.handler_errors)
                    self.handler_errors.append((mtype, repr(exc)))
                    metrics.inc("byzpy_node_handler_errors_total", labels={"type": mtype})
                    del self.handler_errors[:-64]
                    if first:
                        import warnings

                        warnings.warn(f"node {self.node_id!r}: handler of {mtype!r} messages raised {exc!r}; "
                                      "message processing continues (see node.handler_errors)")
        except asyncio.CancelledError:
            pass

    async def handle_incoming_message(self, from_node_id: str, message_type: str, payload: Any) -> None:
        """Deliver one message: wake pipelines waiting for ``message_type``, then call the registered handler."""
        self.scheduler.deliver_message(message_type, payload)
        handler = self._message_handlers.get(message_type)
        if handler is not None:
            await handler(from_node_id, payload)

    def register_message_handler(self, message_type: str,
                                 handler: Callable[[str, Any], Awaitable[None]]) -> None:
        """``await handler(sender_id, payload)`` will be called for every incoming message of ``message_type`` (one handler
        per type; a later registration replaces the earlier one).
        """
        self._message_handlers[message_type] = handler

    def _register_default_handlers(self) -> None:
        pass

    # ---- outbound ---------------------------------------------------------------------------
    def _require_running(self) -> None:
        if not self._running:
            raise RuntimeError("Node not started")

    async def send_message(self, to_node_id: NodeId, message_type: str, payload: Any) -> None:
        """Send to one out-neighbour (``ValueError`` when the topology does not allow it, ``RuntimeError`` before ``start``)."""
        self._require_running()
        if not self.message_router.can_send_to(to_node_id):
            raise ValueError(f"Cannot send to {to_node_id} (not a neighbor)")
        await self.message_router.route_direct(to_node_id, message_type, payload, self.context)

    async def broadcast_message(self, message_type: str, payload: Any) -> None:
        """Send to every out-neighbour; unreachable peers are skipped."""
        self._require_running()
        await self.message_router.route_broadcast(message_type, payload, self.context)

    async def multicast_message(self, to_node_ids: List[NodeId], message_type: str, payload: Any) -> None:
        """Send to the listed nodes, all of which must be out-neighbours."""
        self._require_running()
        await self.message_router.route_multicast(to_node_ids, message_type, payload, self.context)

    def get_neighbors(self) -> List[NodeId]:
        """Ids of the nodes this node may send to."""
        return self.message_router.get_out_neighbors()

    def get_in_neighbors(self) -> List[NodeId]:
        """Ids of the nodes that may send to this node."""
        return self.message_router.get_in_neighbors()

    # ---- pipelines ----------------------------------------------------------------------------
    async def execute_pipeline(self, pipeline_name: str, inputs: Mapping[str, Any], *,
                               triggered_by: Optional[str] = None) -> Dict[str, Any]:
        """Run the application's pipeline ``pipeline_name`` on this node's message-aware scheduler and pool; returns
        ``{output node name: value}``.  Inputs declared as message sources are filled from incoming messages.
        """
        self._require_running()
        pipeline = self.application._pipelines.get(pipeline_name)
        if pipeline is None:
            raise KeyError(f"Unknown pipeline: {pipeline_name}")
        self.scheduler.graph = pipeline.graph
        self.scheduler.metadata["scheduler"] = self.scheduler
        return await self.scheduler.run(inputs)

    async def start_autonomous_task(self, task_coro: Awaitable[Any], name: str = "autonomous_task") -> asyncio.Task:
        """Schedule ``task_coro`` as a named background task of the node (cancelled on shutdown); names are unique."""
        refused = None
        if not self._running:
            refused = RuntimeError("Node must be started before starting autonomous tasks")
        elif name in self._autonomous_tasks:
            refused = ValueError(f"Autonomous task with name '{name}' already exists")
        if refused is not None:
            if asyncio.iscoroutine(task_coro):
                task_coro.close()       # a coroutine object we will never run: do not leave it to warn at GC
            raise refused
        task = asyncio.ensure_future(task_coro)
        self._autonomous_tasks[name] = task
        return task


__all__ = ["DecentralizedNode"]
