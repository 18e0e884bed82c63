"""Per-node application runtime: an ``ActorPool`` plus named pipelines (computation graphs)
executed on demand (reference engine/node/application.py:12-269).

Reserved pipeline names: ``"aggregate"`` and ``"honest_gradient"`` on honest nodes, ``"attack"``
on Byzantine nodes.  ``*_sync`` helpers spin a private event loop and refuse to run from inside
a running loop.
"""
from __future__ import annotations

import asyncio
from dataclasses import dataclass
from typing import Any, Dict, Iterable, Mapping, Optional, Sequence, Union

from ..graph.graph import ComputationGraph
from ..graph.pool import ActorPool, ActorPoolConfig
from ..graph.scheduler import NodeScheduler


@dataclass(frozen=True)
class NodePipeline:
    """A registered pipeline of a :class:`NodeApplication`: the computation graph and the metadata merged into each run.
    """

    graph: ComputationGraph
    metadata: Optional[Mapping[str, Any]] = None


def _run_blocking(coro_fn):
    try:
        asyncio.get_running_loop()
    except RuntimeError:
        return asyncio.run(coro_fn())
    raise RuntimeError("run_pipeline_sync() cannot be called from an async context; "
                       "use the async pipeline APIs instead.")


class NodeApplication:
    """What a node can compute: an :class:`~byzpy_b200.engine.graph.pool.ActorPool` of its own plus named pipelines.

    A pipeline is a :class:`~byzpy_b200.engine.graph.graph.ComputationGraph` registered under a name and run on
    demand on the node's pool.  Decentralized nodes (:class:`~byzpy_b200.engine.node.decentralized.DecentralizedNode`)
    execute their training logic through these pipelines.

    Parameters
    ----------
    name : str
        Node name (appears in scheduler metadata and error messages).
    actor_pool : ActorPool or sequence of ActorPoolConfig
        The node's private pool, or the configuration to build one from.
    metadata : mapping, optional
        Merged into every pipeline run's metadata.

    Notes
    -----
    ``register_pipeline(name, graph, metadata=None)``; ``await run_pipeline(name, inputs)`` returns the graph's outputs
    by node name; ``run_pipeline_sync`` does the same from synchronous code (it refuses to run inside a running event
    loop); ``has_pipeline`` / ``list_pipelines``; ``await shutdown()`` closes the pool.

    Examples
    --------
    >>> import asyncio, torch
    >>> from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
    >>> from byzpy_b200.engine.graph.ops import make_single_operator_graph
    >>> from byzpy_b200.engine.graph.pool import ActorPoolConfig
    >>> from byzpy_b200.engine.node.application import NodeApplication
    >>> app = NodeApplication(name="n0", actor_pool=[ActorPoolConfig("thread", count=1)])
    >>> app.register_pipeline("robust", make_single_operator_graph(node_name="agg", operator=CoordinateWiseMedian(),
    ...                                                              input_keys=("gradients",)))
    >>> async def demo():
    ...     try:
    ...         return await app.run_pipeline("robust", {"gradients": [torch.tensor([1.0]), torch.tensor([9.0]), torch.tensor([2.0])]})
    ...     finally:
    ...         await app.shutdown()
    >>> asyncio.run(demo())
    {'agg': tensor([2.])}
    """

    def __init__(self, *, name: str, actor_pool: Union[ActorPool, Sequence[ActorPoolConfig]],
                 metadata: Optional[Mapping[str, Any]] = None) -> None:
        self.name = name
        self._pool = actor_pool if isinstance(actor_pool, ActorPool) else ActorPool(actor_pool)
        self._pipelines: Dict[str, NodePipeline] = {}
        self._base_metadata: Dict[str, Any] = dict(metadata or {})

    @property
    def pool(self) -> ActorPool:
        """The node's actor pool."""
        return self._pool

    def register_pipeline(self, name: str, graph: ComputationGraph, *,
                          metadata: Optional[Mapping[str, Any]] = None) -> None:
        """Register ``graph`` under ``name`` (``ValueError`` when the name is taken); ``metadata`` is merged into each run."""
        if name in self._pipelines:
            raise ValueError(f"Pipeline {name!r} already registered for node {self.name!r}.")
        self._pipelines[name] = NodePipeline(graph=graph, metadata=dict(metadata or {}))

    def has_pipeline(self, name: str) -> bool:
        """Whether a pipeline called ``name`` is registered."""
        return name in self._pipelines

    def list_pipelines(self) -> Iterable[str]:
        """Names of the registered pipelines."""
        return self._pipelines.keys()

    def _pipeline(self, name: str) -> NodePipeline:
        try:
            return self._pipelines[name]
        except KeyError:
            raise KeyError(f"Unknown pipeline {name!r} for node {self.name!r}.") from None

    async def run_pipeline(self, name: str, inputs: Mapping[str, Any], *,
                           metadata: Optional[Mapping[str, Any]] = None) -> Dict[str, Any]:
        """Run pipeline ``name`` on the node's pool; returns ``{output node name: value}`` (``KeyError`` for an unknown name)."""
        pipe = self._pipeline(name)
        meta: Dict[str, Any] = {"node": self.name, "pipeline": name}
        meta.update(self._base_metadata)
        meta.update(pipe.metadata or {})
        meta.update(metadata or {})
        return await NodeScheduler(pipe.graph, pool=self._pool, metadata=meta).run(inputs)

    def run_pipeline_sync(self, name: str, inputs: Mapping[str, Any], *,
                          metadata: Optional[Mapping[str, Any]] = None) -> Dict[str, Any]:
        """:meth:`run_pipeline` for synchronous callers; raises ``RuntimeError`` inside a running event loop."""
        return _run_blocking(lambda: self.run_pipeline(name, inputs, metadata=metadata))

    async def _single(self, name: str, inputs, metadata, missing_msg: str) -> Any:
        if not self.has_pipeline(name):
            raise KeyError(missing_msg)
        return next(iter((await self.run_pipeline(name, inputs, metadata=metadata)).values()))

    def _single_sync(self, name: str, inputs, metadata, missing_msg: str) -> Any:
        if not self.has_pipeline(name):
            raise KeyError(missing_msg)
        return next(iter(self.run_pipeline_sync(name, inputs, metadata=metadata).values()))

    async def shutdown(self) -> None:
        """Shut the node's pool down."""
        await self._pool.shutdown()


class HonestNodeApplication(NodeApplication):
    """A :class:`NodeApplication` with the two pipelines an honest node is expected to have.

    ``"aggregate"`` (input ``gradients``) is run by :meth:`aggregate`; ``"honest_gradient"`` by :meth:`honest_gradient`.
    Both return the single output of the pipeline; the ``*_sync`` variants are for synchronous callers.  Calling one
    whose pipeline was never registered raises ``KeyError``.
    """

    AGGREGATION_PIPELINE = "aggregate"
    GRADIENT_PIPELINE = "honest_gradient"

    async def aggregate(self, *, gradients: Sequence[Any], metadata=None) -> Any:
        """Result of the ``"aggregate"`` pipeline on ``gradients``."""
        return await self._single(self.AGGREGATION_PIPELINE, {"gradients": gradients}, metadata,
                                  f"No aggregation pipeline registered on node {self.name!r}.")

    def aggregate_sync(self, *, gradients: Sequence[Any], metadata=None) -> Any:
        """Synchronous :meth:`aggregate`."""
        return self._single_sync(self.AGGREGATION_PIPELINE, {"gradients": gradients}, metadata,
                                 f"No aggregation pipeline registered on node {self.name!r}.")

    async def honest_gradient(self, inputs: Mapping[str, Any], *, metadata=None) -> Any:
        """Result of the ``"honest_gradient"`` pipeline on ``inputs`` (usually ``{"x": ..., "y": ...}``)."""
        return await self._single(self.GRADIENT_PIPELINE, inputs, metadata,
                                  f"No honest gradient pipeline registered on node {self.name!r}.")

    def honest_gradient_sync(self, inputs: Mapping[str, Any], *, metadata=None) -> Any:
        """Synchronous :meth:`honest_gradient`."""
        return self._single_sync(self.GRADIENT_PIPELINE, inputs, metadata,
                                 f"No honest gradient pipeline registered on node {self.name!r}.")


class ByzantineNodeApplication(NodeApplication):
    """A :class:`NodeApplication` with the ``"attack"`` pipeline a Byzantine node is expected to have, run by
    :meth:`run_attack` / :meth:`run_attack_sync` with whatever inputs the attack operator reads (``honest_grads``,
    ``base_grad``, ``model`` / ``x`` / ``y``).
    """

    ATTACK_PIPELINE = "attack"

    async def run_attack(self, *, inputs: Mapping[str, Any], metadata=None) -> Any:
        """Result of the ``"attack"`` pipeline on ``inputs``."""
        return await self._single(self.ATTACK_PIPELINE, inputs, metadata,
                                  f"No attack pipeline registered on node {self.name!r}.")

    def run_attack_sync(self, *, inputs: Mapping[str, Any], metadata=None) -> Any:
        """Synchronous :meth:`run_attack`."""
        return self._single_sync(self.ATTACK_PIPELINE, inputs, metadata,
                                 f"No attack pipeline registered on node {self.name!r}.")


__all__ = ["NodeApplication", "NodePipeline", "HonestNodeApplication", "ByzantineNodeApplication"]
