"""Decentralized peer-to-peer training over ``DecentralizedNode`` s
(reference engine/peer_to_peer/runner.py:184-392).

A round (per honest node i):  local half step ``theta_i <- theta_i - lr * grad``  ->  broadcast
``theta_i^{t+1/2}`` to the out-neighbours  ->  robustly aggregate own + received vectors  ->
**write the aggregate back into the model** (the reference computes it with a hard-coded
median and discards it; the intended semantics are ``P2PHonestMixin.
p2p_aggregate_and_set``, reference engine/node/mixin.py:71-80, which is what runs here, with the
node's own ``p2p_agg`` / ``p2p_pre``).  Byzantine nodes wait for their honest in-neighbours'
vectors, run ``p2p_broadcast_vector`` on them and broadcast the result.

The round is message-driven: a node proceeds when it has one vector from every (unique)
in-neighbour (``recv_timeout`` bounds the wait) -- no fixed sleeps.  Node objects are reached
directly when they live in this process (thread/gpu actor backends), otherwise through their
actor proxy.  The default context is ``InProcessContext``; pass ``context_factory`` for
``ProcessContext`` / ``RemoteContext`` / ``MeshRemoteContext`` deployments, or set
``BYZPY_P2P_CONTEXT=process`` to get the reference's default (one OS process per node,
reference peer_to_peer/runner.py:221-224) without touching the call sites.
"""
from __future__ import annotations

import asyncio
import os
import inspect
from typing import Any, Callable, Dict, List, Optional

import torch

from ...aggregators.coordinate_wise.median import CoordinateWiseMedian
from ...utils import metrics
from ..graph.ops import CallableOp, make_single_operator_graph
from ..graph.pool import ActorPoolConfig
from ..node.application import ByzantineNodeApplication, HonestNodeApplication
from ..node.cluster import DecentralizedCluster
from ..node.context import InProcessContext, NodeContext, ProcessContext
from ..node.decentralized import DecentralizedNode
from .topology import Topology

# process-wide registries (also filled inside ProcessContext children)
_NODE_OBJECT_REGISTRY: Dict[str, Any] = {}
_ACTOR_REGISTRY: Dict[str, Any] = {}


def _local_object(actor: Any) -> Any:
    """The raw node object when the actor lives in this process, else None."""
    backend = getattr(getattr(actor, "_ref", None), "_backend", None)
    obj = getattr(backend, "_obj", None)
    if obj is not None:
        return obj
    if hasattr(actor, "p2p_half_step") or hasattr(actor, "p2p_broadcast_vector"):
        return actor if not hasattr(actor, "_ref") else None
    return None


async def _invoke(node_key: str, method: str, *args, **kwargs) -> Any:
    target = _NODE_OBJECT_REGISTRY.get(node_key)
    if target is None:
        target = _ACTOR_REGISTRY.get(node_key)
    if target is None:
        raise RuntimeError(f"Node object {node_key} not found in registry")
    out = getattr(target, method)(*args, **kwargs)
    if inspect.isawaitable(out):
        out = await out
    return out


def _local_node(node_key: str) -> Any:
    """Registry lookup behind a module-level function: the pipeline closures are cloudpickled by value into
    ``ProcessContext`` children, and a closure naming the registry dict itself would drag every node object
    along (module-level functions travel by reference)."""
    return _NODE_OBJECT_REGISTRY.get(node_key)


def _register(node_key: str, actor: Any) -> None:
    obj = _local_object(actor)
    if obj is not None:
        _NODE_OBJECT_REGISTRY[node_key] = obj
    _ACTOR_REGISTRY[node_key] = actor


def _create_honest_node_application(actor: Any, node_id: str, lr: float) -> HonestNodeApplication:
    app = HonestNodeApplication(name=f"honest_{node_id}",
                                actor_pool=[ActorPoolConfig(backend="thread", count=1)])
    key = f"honest_{node_id}"
    _register(key, actor)

    async def half_step(lr: float):
        return await _invoke(key, "p2p_half_step", lr)

    async def aggregate_and_set(own, received):
        """Robust aggregation of own + received, written back into the model.  Nodes built on
        ``P2PHonestMixin`` provide ``p2p_aggregate_and_set``; duck-typed nodes may only offer the older
        ``p2p_aggregate(vectors)`` hook, and a node with neither gets the coordinate-wise median of the
        candidates loaded through ``set_param_vector`` when it has one (the reference computes exactly that
        median and then drops it)."""
        received = list(received)
        local = _local_node(key)
        if local is None or hasattr(local, "p2p_aggregate_and_set"):
            try:
                await _invoke(key, "p2p_aggregate_and_set", own, received)
                return True
            except AttributeError:
                if local is not None:
                    raise                      # raised inside the node's own method: not a missing hook
        if local is None or hasattr(local, "p2p_aggregate"):
            try:
                return await _invoke(key, "p2p_aggregate", [own, *received])
            except AttributeError:
                if local is not None:
                    raise
        agg = CoordinateWiseMedian().aggregate([own, *received])
        if local is None or hasattr(local, "set_param_vector"):
            try:
                await _invoke(key, "set_param_vector", agg)
            except AttributeError:
                if local is not None:
                    raise
        return agg

    app.register_pipeline("half_step", make_single_operator_graph(
        node_name="half_step", operator=CallableOp(half_step, input_mapping={"lr": "lr"}),
        input_keys=("lr",)))
    app.register_pipeline("aggregate_and_set", make_single_operator_graph(
        node_name="aggregate_and_set",
        operator=CallableOp(aggregate_and_set, input_mapping={"own": "own", "received": "received"}),
        input_keys=("own", "received")))
    # kept for API parity: a plain aggregation pipeline (median), unused by run_round_async
    app.register_pipeline("aggregate", make_single_operator_graph(
        node_name="aggregate", operator=CoordinateWiseMedian(), input_keys=("gradients",)))
    return app


def _create_byzantine_node_application(actor: Any, node_id: str) -> ByzantineNodeApplication:
    app = ByzantineNodeApplication(name=f"byz_{node_id}",
                                   actor_pool=[ActorPoolConfig(backend="thread", count=1)])
    key = f"byz_{node_id}"
    _register(key, actor)

    async def broadcast_vector(neighbor_vectors: Optional[List[torch.Tensor]] = None,
                               like: Optional[torch.Tensor] = None):
        return await _invoke(key, "p2p_broadcast_vector", neighbor_vectors=neighbor_vectors, like=like)

    app.register_pipeline("broadcast", make_single_operator_graph(
        node_name="broadcast",
        operator=CallableOp(broadcast_vector, input_mapping={"neighbor_vectors": "neighbor_vectors",
                                                             "like": "like"}),
        input_keys=("neighbor_vectors", "like")))
    return app


def _default_context() -> NodeContext:
    kind = os.environ.get("BYZPY_P2P_CONTEXT", "inprocess").strip().lower()
    if kind in ("process", "processcontext"):
        return ProcessContext()
    if kind in ("", "inprocess", "inprocesscontext"):
        return InProcessContext()
    raise ValueError(f"BYZPY_P2P_CONTEXT={kind!r}: expected 'inprocess' or 'process'")


class DecentralizedPeerToPeer:
    """Message-driven gossip training over :class:`~byzpy_b200.engine.node.decentralized.DecentralizedNode` objects.

    Every node (honest first, then Byzantine; ids ``"0" .. "n-1"`` in that order, matching the topology indices) gets a
    decentralized node with ``half_step`` / ``aggregate`` (honest) or ``broadcast`` (Byzantine) pipelines wrapped around
    the user's node object.  A round: honest nodes take a local half step and send the resulting parameter vector to
    their out-neighbours; Byzantine nodes wait for their honest in-neighbours' vectors, craft one vector from them and
    broadcast it; every honest node robustly aggregates its own and the received vectors and loads the result.

    Parameters
    ----------
    honest_nodes, byzantine_nodes : list
        Node objects or node actors implementing the P2P mixin methods (``p2p_half_step`` + ``p2p_aggregate_and_set`` or
        just ``p2p_aggregate``; ``p2p_broadcast_vector``).
    topology : Topology
    lr : float, default 0.05
    context_factory : callable, optional
        ``(node_id, index) -> NodeContext``; default :class:`~byzpy_b200.engine.node.context.InProcessContext` (or
        ``ProcessContext`` with ``BYZPY_P2P_CONTEXT=process``).
    recv_timeout : float, default 30.0
        Longest a node waits for its neighbours' vectors before going on with what it has.

    Notes
    -----
    ``await start()``, ``await run_round_async()`` (``run_round()`` from synchronous code), ``await stop()``;
    ``rounds`` counts completed rounds, ``cluster`` is the underlying
    :class:`~byzpy_b200.engine.node.cluster.DecentralizedCluster`.  Usually reached through
    :class:`~byzpy_b200.engine.peer_to_peer.train.PeerToPeer`.
    """

    def __init__(self, honest_nodes: List[Any], byzantine_nodes: List[Any], topology: Topology, *,
                 lr: float = 0.05, context_factory: Optional[Callable[[str, int], NodeContext]] = None,
                 recv_timeout: float = 30.0) -> None:
        self._cluster = DecentralizedCluster()
        self.topology = topology
        self.honest = list(honest_nodes)
        self.byz = list(byzantine_nodes or [])
        self.lr = lr
        self.context_factory = context_factory
        self.recv_timeout = recv_timeout
        self._node_applications: Dict[str, Any] = {}
        self._decentralized_nodes: Dict[str, DecentralizedNode] = {}
        self._gradient_cache: Dict[str, List[torch.Tensor]] = {}
        self._arrived: Dict[str, asyncio.Event] = {}
        self.rounds = 0

    @property
    def cluster(self) -> DecentralizedCluster:
        return self._cluster

    def _n(self) -> int:
        return len(self.honest) + len(self.byz)

    async def start(self) -> None:
        for idx in range(self._n()):
            node_id = str(idx)
            context = self.context_factory(node_id, idx) if self.context_factory else _default_context()
            if idx < len(self.honest):
                actor = self.honest[idx]
                try:
                    actor.lr = self.lr
                except Exception:
                    pass
                app = _create_honest_node_application(actor, node_id, self.lr)
            else:
                app = _create_byzantine_node_application(self.byz[idx - len(self.honest)], node_id)
            self._node_applications[node_id] = app
            node = await self._cluster.add_node(node_id=node_id, application=app,
                                                topology=self.topology, context=context)
            self._decentralized_nodes[node_id] = node
            self._gradient_cache[node_id] = []
            self._arrived[node_id] = asyncio.Event()

            def make_handler(nid: str):
                async def on_vector(from_id, payload):
                    vec = payload.get("vector") if isinstance(payload, dict) else payload
                    self._gradient_cache.setdefault(nid, []).append(vec)
                    self._arrived[nid].set()
                return on_vector

            node.register_message_handler("gradient", make_handler(node_id))
        await self._cluster.start_all()

    async def stop(self) -> None:
        await self._cluster.shutdown_all()
        for idx in range(self._n()):
            key = f"{'honest' if idx < len(self.honest) else 'byz'}_{idx}"
            _NODE_OBJECT_REGISTRY.pop(key, None)
            _ACTOR_REGISTRY.pop(key, None)

    async def _wait_for(self, node_id: str, count: int) -> List[torch.Tensor]:
        """Wait until ``count`` vectors arrived for ``node_id`` (or the timeout hits)."""
        loop = asyncio.get_running_loop()
        deadline = loop.time() + self.recv_timeout
        while len(self._gradient_cache.get(node_id, [])) < count:
            remaining = deadline - loop.time()
            if remaining <= 0:
                break
            ev = self._arrived[node_id]
            ev.clear()
            if len(self._gradient_cache.get(node_id, [])) >= count:
                break
            try:
                await asyncio.wait_for(ev.wait(), timeout=remaining)
            except asyncio.TimeoutError:
                break
        return list(self._gradient_cache.get(node_id, []))

    def _in_count(self, idx: int, honest_only: bool = False) -> int:
        ins = list(dict.fromkeys(self.topology.in_.get(idx, [])))
        if honest_only:
            ins = [j for j in ins if j < len(self.honest)]
        return len(ins)

    async def run_round_async(self) -> None:
        n_h = len(self.honest)
        # 1) honest half steps (concurrently)
        async def half(i: int):
            node = self._decentralized_nodes[str(i)]
            return (await node.execute_pipeline("half_step", {"lr": self.lr}))["half_step"]

        halves = await asyncio.gather(*[half(i) for i in range(n_h)])
        half_step_results = {str(i): halves[i] for i in range(n_h)}
        template = halves[0] if halves else None
        # 2) honest broadcasts
        for i in range(n_h):
            await self._decentralized_nodes[str(i)].broadcast_message(
                "gradient", {"vector": half_step_results[str(i)]})
        # 3) Byzantine nodes: wait for the honest in-neighbours, attack, broadcast
        for j in range(n_h, self._n()):
            node_id = str(j)
            node = self._decentralized_nodes[node_id]
            seen = await self._wait_for(node_id, self._in_count(j, honest_only=True))
            like = template if template is not None else (seen[0] if seen else None)
            if like is None:
                continue
            out = await node.execute_pipeline("broadcast", {"neighbor_vectors": seen or [like], "like": like})
            self._gradient_cache[node_id] = []
            await node.broadcast_message("gradient", {"vector": out["broadcast"]})
        # 4) honest aggregation, written back into the models
        async def finish(i: int):
            node_id = str(i)
            received = await self._wait_for(node_id, self._in_count(i))
            self._gradient_cache[node_id] = []
            if received:
                await self._decentralized_nodes[node_id].execute_pipeline(
                    "aggregate_and_set", {"own": half_step_results[node_id], "received": received})

        await asyncio.gather(*[finish(i) for i in range(n_h)])
        for j in range(n_h, self._n()):
            self._gradient_cache[str(j)] = []
        self.rounds += 1
        metrics.inc("byzpy_p2p_rounds_total")

    def run_round(self) -> None:
        asyncio.run(self.run_round_async())


__all__ = ["DecentralizedPeerToPeer"]
