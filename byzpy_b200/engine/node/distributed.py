"""User-facing node bases whose heavy work runs through pipelines on an actor pool
(reference engine/node/distributed.py:19-320).

``DistributedHonestNode``: subclasses implement ``local_honest_gradient(x=, y=)`` (and the
``next_batch`` / ``apply_server_gradient`` node contract); aggregation and gradient pipelines are
registered automatically.  ``DistributedByzantineNode``: wraps an :class:`Attack` (inputs picked
from its ``uses_*`` flags) or, when the subclass overrides ``byzantine_gradient``, ships that
override to the pool through a ``RemoteCallableOp``.  Unlike the reference, ``aggregate`` hands
tensors to the aggregator directly (no host shm round trip; aggregators accept handles too).
"""
from __future__ import annotations

import inspect
from typing import Mapping, Optional, Sequence, Union

import torch

from ...aggregators.base import Aggregator
from ...attacks.base import Attack
from ..graph.ops import CallableOp, RemoteCallableOp, make_single_operator_graph
from ..graph.pool import ActorPool, ActorPoolConfig
from .application import ByzantineNodeApplication, HonestNodeApplication, NodeApplication
from .base import ByzantineNode, HonestNode

PoolSpec = Union[ActorPool, Sequence[ActorPoolConfig]]


class _DistributedNodeBase:
    def __init__(self, *, app_cls: type, name: Optional[str], actor_pool: PoolSpec,
                 metadata: Optional[Mapping[str, object]] = None) -> None:
        self._app: NodeApplication = app_cls(name=name or type(self).__name__, actor_pool=actor_pool,
                                             metadata=metadata)

    @property
    def application(self) -> NodeApplication:
        return self._app

    @property
    def pool(self) -> ActorPool:
        return self._app.pool

    async def shutdown_distributed(self) -> None:
        await self._app.shutdown()


class DistributedHonestNode(_DistributedNodeBase, HonestNode):
    """Base class of honest nodes whose gradient and aggregation work runs as pipelines on the node's own actor pool.

    Subclass it, implement ``local_honest_gradient(*, x, y)`` plus ``next_batch`` / ``apply_server_gradient``; the
    constructor registers the ``"honest_gradient"`` and ``"aggregate"`` pipelines.  ``honest_gradient(x, y)`` then runs
    the gradient pipeline, ``await aggregate(gradients)`` / ``aggregate_sync`` the node's aggregator.

    Parameters
    ----------
    actor_pool : ActorPool or sequence of ActorPoolConfig
    aggregator : Aggregator
        Used when this node aggregates (P2P) -- the parameter server has its own.
    metadata : mapping, optional
    name : str, optional

    Notes
    -----
    ``examples/ps/nodes.py`` (``DistributedPSHonestNode``) is the canonical subclass; ``await shutdown_distributed()``
    closes the pool.
    """

    def __init__(self, *, actor_pool: PoolSpec, aggregator: Aggregator,
                 metadata: Optional[Mapping[str, object]] = None, name: Optional[str] = None) -> None:
        self._aggregator = aggregator
        super().__init__(app_cls=HonestNodeApplication, name=name, actor_pool=actor_pool,
                         metadata=metadata)
        app = self._app
        app.register_pipeline(app.AGGREGATION_PIPELINE, make_single_operator_graph(
            node_name="aggregate", operator=aggregator, input_keys=("gradients",)))
        grad_op = CallableOp(self._gradient_callable, input_mapping={"x": "x", "y": "y"})
        app.register_pipeline(app.GRADIENT_PIPELINE, make_single_operator_graph(
            node_name="honest_gradient", operator=grad_op, input_keys=("x", "y")))

    def local_honest_gradient(self, *, x, y):
        raise NotImplementedError("local_honest_gradient() must be implemented by subclasses.")

    def _gradient_callable(self, *, x, y):
        return self.local_honest_gradient(x=x, y=y)

    async def aggregate(self, gradients) -> torch.Tensor:
        return await self._app.aggregate(gradients=list(gradients))

    def aggregate_sync(self, gradients):
        return self._app.aggregate_sync(gradients=list(gradients))

    def honest_gradient(self, x, y):
        return self._app.honest_gradient_sync({"x": x, "y": y})


class DistributedByzantineNode(_DistributedNodeBase, ByzantineNode):
    """Base class of Byzantine nodes whose attack runs as the ``"attack"`` pipeline on the node's own actor pool.

    Either pass an :class:`~byzpy_b200.attacks.base.Attack` (its ``uses_*`` flags decide which of ``honest_grads`` /
    ``base_grad`` / ``model, x, y`` it is given), or override ``byzantine_gradient`` in the subclass -- the override is
    shipped to the pool as the body of the pipeline and the public ``byzantine_gradient(x, y, honest_grads)`` keeps the
    calling convention orchestrators use.

    Parameters
    ----------
    actor_pool : ActorPool or sequence of ActorPoolConfig
    attack : Attack, optional
    metadata : mapping, optional
    name : str, optional
    """

    _distributed_user_bz = None

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        own = cls.__dict__.get("byzantine_gradient")
        if own is not None and own is not DistributedByzantineNode.byzantine_gradient:
            # the user's implementation (any signature) becomes the body of the attack pipeline; the public
            # ``byzantine_gradient(x, y, honest_grads)`` entry point goes back to the dispatcher below, so
            # the parameter server's calling convention keeps working and the call runs on the node's pool
            cls._distributed_user_bz = own
            cls.byzantine_gradient = DistributedByzantineNode.byzantine_gradient
        else:
            cls._distributed_user_bz = getattr(cls, "_distributed_user_bz", None)

    def __init__(self, *, actor_pool: PoolSpec, attack: Optional[Attack] = None,
                 metadata: Optional[Mapping[str, object]] = None, name: Optional[str] = None) -> None:
        self.attack = attack
        super().__init__(app_cls=ByzantineNodeApplication, name=name, actor_pool=actor_pool,
                         metadata=metadata)
        self._custom_bz_callable = None
        self._custom_input_keys: tuple = ()
        self._custom_required_keys: tuple = ()
        user_impl = getattr(type(self), "_distributed_user_bz", None)
        app = self._app
        if user_impl is not None:
            self._custom_bz_callable = user_impl.__get__(self, type(self))
            keys, required, defaults = [], [], {}
            for pname, param in inspect.signature(self._custom_bz_callable).parameters.items():
                if pname == "self":
                    continue
                keys.append(pname)
                if param.default is inspect.Signature.empty:
                    required.append(pname)
                else:
                    defaults[pname] = param.default
            self._custom_input_keys, self._custom_required_keys = tuple(keys), tuple(required)
            self._custom_defaults = defaults
            op = RemoteCallableOp(self._custom_bz_callable, input_mapping={k: k for k in keys})
            app.register_pipeline(app.ATTACK_PIPELINE, make_single_operator_graph(
                node_name="attack", operator=op, input_keys=keys))
        else:
            if attack is None:
                raise ValueError("DistributedByzantineNode requires an Attack instance when "
                                 "byzantine_gradient is not overridden.")
            app.register_pipeline(app.ATTACK_PIPELINE, make_single_operator_graph(
                node_name="attack", operator=attack, input_keys=self._attack_keys()))

    def _attack_keys(self) -> list:
        keys = []
        if getattr(self.attack, "uses_model_batch", False):
            keys += ["model", "x", "y"]
        if getattr(self.attack, "uses_honest_grads", False):
            keys.append("honest_grads")
        if getattr(self.attack, "uses_base_grad", False):
            keys.append("base_grad")
        return keys

    def prepare_attack_inputs(self, *, x=None, y=None, honest_grads=None, base_grad=None,
                              model=None) -> Mapping[str, object]:
        """Default routing of attack inputs; subclasses may override (e.g. to compute their own
        ``base_grad``)."""
        if self._custom_bz_callable is not None:
            raise RuntimeError("prepare_attack_inputs should not be used when byzantine_gradient "
                               "is overridden.")
        out: dict = {}
        if getattr(self.attack, "uses_model_batch", False):
            if model is None:
                raise ValueError("Attack requires 'model' but none was supplied.")
            if x is None or y is None:
                raise ValueError("Attack requires batch inputs 'x' and 'y'.")
            out.update(model=model, x=x, y=y)
        if getattr(self.attack, "uses_honest_grads", False):
            if honest_grads is None:
                raise ValueError("Attack requires 'honest_grads'.")
            out["honest_grads"] = honest_grads
        if getattr(self.attack, "uses_base_grad", False):
            if base_grad is None:
                raise ValueError("Attack requires 'base_grad'.")
            out["base_grad"] = base_grad
        return out

    def _build_custom_inputs(self, **data) -> Mapping[str, object]:
        out: dict = {}
        for key in self._custom_input_keys:
            value = data.get(key)
            if value is None and key in self._custom_required_keys:
                raise ValueError(f"Custom byzantine_gradient requires argument {key!r}.")
            if value is not None:
                out[key] = value
        return out

    def _complete_custom_inputs(self, inputs: Mapping[str, object]) -> Mapping[str, object]:
        """The attack pipeline's graph names every parameter of the user's function as an input; optional
        ones the caller left out are bound to their declared defaults."""
        out = dict(inputs)
        for key, default in getattr(self, "_custom_defaults", {}).items():
            out.setdefault(key, default)
        return out

    def _inputs(self, **data) -> Mapping[str, object]:
        if self._custom_bz_callable is not None:
            return self._complete_custom_inputs(self._build_custom_inputs(**data))
        return self.prepare_attack_inputs(**data)

    async def run_attack(self, *, inputs: Mapping[str, object]) -> torch.Tensor:
        return await self._app.run_attack(inputs=inputs)

    def byzantine_gradient(self, x, y, honest_grads=None):
        return self._app.run_attack_sync(inputs=self._inputs(x=x, y=y, honest_grads=honest_grads))

    async def byzantine_gradient_async(self, *, x=None, y=None, honest_grads=None, base_grad=None,
                                       model=None):
        return await self._app.run_attack(inputs=self._inputs(
            x=x, y=y, honest_grads=honest_grads, base_grad=base_grad, model=model))


__all__ = ["DistributedHonestNode", "DistributedByzantineNode"]
