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
    def __init__(self, *, name: str, actor_pool: Union[ActorPool, Sequence[ActorPoolConfig]],
                 metadata: Optional[Mapping[str, Any]] = None) -> None:
        self.name = name
        self._pool = actor_pool if isinstance(actor_pool, ActorPool) else ActorPool(actor_pool)
        self._pipelines: Dict[str, NodePipeline] = {}
        self._base_metadata: Dict[str, Any] = dict(metadata or {})

    @property
    def pool(self) -> ActorPool:
        return self._pool

    def register_pipeline(self, name: str, graph: ComputationGraph, *,
                          metadata: Optional[Mapping[str, Any]] = None) -> None:
        if name in self._pipelines:
            raise ValueError(f"Pipeline {name!r} already registered for node {self.name!r}.")
        self._pipelines[name] = NodePipeline(graph=graph, metadata=dict(metadata or {}))

    def has_pipeline(self, name: str) -> bool:
        return name in self._pipelines

    def list_pipelines(self) -> Iterable[str]:
        return self._pipelines.keys()

    def _pipeline(self, name: str) -> NodePipeline:
        try:
            return self._pipelines[name]
        except KeyError:
            raise KeyError(f"Unknown pipeline {name!r} for node {self.name!r}.") from None

    async def run_pipeline(self, name: str, inputs: Mapping[str, Any], *,
                           metadata: Optional[Mapping[str, Any]] = None) -> Dict[str, Any]:
        pipe = self._pipeline(name)
        meta: Dict[str, Any] = {"node": self.name, "pipeline": name}
        meta.update(self._base_metadata)
        meta.update(pipe.metadata or {})
        meta.update(metadata or {})
        return await NodeScheduler(pipe.graph, pool=self._pool, metadata=meta).run(inputs)

    def run_pipeline_sync(self, name: str, inputs: Mapping[str, Any], *,
                          metadata: Optional[Mapping[str, Any]] = None) -> Dict[str, Any]:
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
        await self._pool.shutdown()


class HonestNodeApplication(NodeApplication):
    AGGREGATION_PIPELINE = "aggregate"
    GRADIENT_PIPELINE = "honest_gradient"

    async def aggregate(self, *, gradients: Sequence[Any], metadata=None) -> Any:
        return await self._single(self.AGGREGATION_PIPELINE, {"gradients": gradients}, metadata,
                                  f"No aggregation pipeline registered on node {self.name!r}.")

    def aggregate_sync(self, *, gradients: Sequence[Any], metadata=None) -> Any:
        return self._single_sync(self.AGGREGATION_PIPELINE, {"gradients": gradients}, metadata,
                                 f"No aggregation pipeline registered on node {self.name!r}.")

    async def honest_gradient(self, inputs: Mapping[str, Any], *, metadata=None) -> Any:
        return await self._single(self.GRADIENT_PIPELINE, inputs, metadata,
                                  f"No honest gradient pipeline registered on node {self.name!r}.")

    def honest_gradient_sync(self, inputs: Mapping[str, Any], *, metadata=None) -> Any:
        return self._single_sync(self.GRADIENT_PIPELINE, inputs, metadata,
                                 f"No honest gradient pipeline registered on node {self.name!r}.")


class ByzantineNodeApplication(NodeApplication):
    ATTACK_PIPELINE = "attack"

    async def run_attack(self, *, inputs: Mapping[str, Any], metadata=None) -> Any:
        return await self._single(self.ATTACK_PIPELINE, inputs, metadata,
                                  f"No attack pipeline registered on node {self.name!r}.")

    def run_attack_sync(self, *, inputs: Mapping[str, Any], metadata=None) -> Any:
        return self._single_sync(self.ATTACK_PIPELINE, inputs, metadata,
                                 f"No attack pipeline registered on node {self.name!r}.")


__all__ = ["NodeApplication", "NodePipeline", "HonestNodeApplication", "ByzantineNodeApplication"]
