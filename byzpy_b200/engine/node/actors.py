"""Node actors: spawn a node class on any actor backend and talk to it through an async proxy
(reference engine/node/actors.py:41-91).  The class travels by value (cloudpickle)."""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple, Union

from ..actor.base import ActorBackend, ActorRef
from ..actor.factory import resolve_backend


class NodeActor:
    """A node object hosted by an actor backend, seen through an async proxy.

    ``await NodeActor.spawn(NodeClass, backend=..., args=..., kwargs=...)`` constructs ``NodeClass(*args, **kwargs)``
    *inside* the backend (the class is shipped by value, so the remote side needs this package but not your module) and
    returns the proxy; every attribute access on the proxy is an awaitable call of the same-named method
    (``await actor.honest_gradient_for_next_batch()``).  ``await actor.close()`` stops the backend.

    Parameters of ``spawn``
    -----------------------
    node_cls : type
        The node class.
    backend : str or ActorBackend, default "thread"
        ``"thread"``, ``"process"``, ``"gpu"`` / ``"gpu:<index>"``, ``"tcp://host:port"``, ``"ucx://host:port"``, or a
        backend instance (``configs.actor.set_actor`` builds one from the same strings).
    args, kwargs :
        Constructor arguments of ``node_cls``.

    Examples
    --------
    >>> import asyncio
    >>> from byzpy_b200.engine.node.actors import NodeActor
    >>> class Counter:
    ...     def __init__(self, start=0):
    ...         self.n = start
    ...     def bump(self, by=1):
    ...         self.n += by
    ...         return self.n
    >>> async def demo():
    ...     actor = await NodeActor.spawn(Counter, backend="thread", kwargs={"start": 40})
    ...     try:
    ...         return await actor.bump(2)
    ...     finally:
    ...         await actor.close()
    >>> asyncio.run(demo())
    42
    """

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
    """:class:`NodeActor` for honest nodes (default backend ``"thread"``); orchestrators take a list of these as
    ``honest_nodes``.
    """

    pass


class ByzantineNodeActor(NodeActor):
    """:class:`NodeActor` for Byzantine nodes.  The default backend is ``"process"``, as in the reference: adversarial
    code gets its own interpreter unless the caller asks otherwise.
    """

    @classmethod
    async def spawn(cls, node_cls: Any, *, backend: Union[str, ActorBackend] = "process",
                    args: Tuple[Any, ...] = (), kwargs: Optional[Dict[str, Any]] = None):
        return await super().spawn(node_cls, backend=backend, args=args, kwargs=kwargs)


__all__ = ["NodeActor", "HonestNodeActor", "ByzantineNodeActor"]
