"""Node actors: spawn a node class on any actor backend and talk to it through an async proxy
(reference engine/node/actors.py:41-91).  The class travels by value (cloudpickle)."""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple, Union

from ..actor.base import ActorBackend, ActorRef
from ..actor.factory import resolve_backend


class NodeActor:
    def __init__(self, ref: ActorRef) -> None:
        self._ref = ref

    def __getattr__(self, name: str):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return getattr(self._ref, name)

    @classmethod
    async def spawn(cls, node_cls: Any, *, backend: Union[str, ActorBackend] = "thread",
                    args: Tuple[Any, ...] = (), kwargs: Optional[Dict[str, Any]] = None):
        be = resolve_backend(backend)
        await be.start()
        await be.construct(node_cls, args=tuple(args), kwargs=dict(kwargs or {}))
        return cls(ActorRef(be))

    async def close(self) -> None:
        await self._ref._backend.close()


class HonestNodeActor(NodeActor):
    pass


class ByzantineNodeActor(NodeActor):
    @classmethod
    async def spawn(cls, node_cls: Any, *, backend: Union[str, ActorBackend] = "process",
                    args: Tuple[Any, ...] = (), kwargs: Optional[Dict[str, Any]] = None):
        return await super().spawn(node_cls, backend=backend, args=args, kwargs=kwargs)


__all__ = ["NodeActor", "HonestNodeActor", "ByzantineNodeActor"]
