"""One-operator convenience runner (reference engine/graph/executor.py:33-294).

``OperatorExecutor(op, input_keys=None, pool_config=None, node_name=None)`` builds the
single-node graph + ``NodeScheduler`` (and lazily owns an ``ActorPool`` when ``pool_config`` is
given); ``run_operator`` is the one-shot form.  Input keys are auto-detected for aggregators and
pre-aggregators; attacks need explicit ``input_keys``.  A custom single key different from the
operator's ``input_key`` is transparently renamed.
"""
from __future__ import annotations

from typing import Any, Mapping, Optional, Sequence, Union

from .operator import OpContext, Operator
import weakref

from .ops import make_single_operator_graph
from .pool import ActorPool, ActorPoolConfig
from .scheduler import NodeScheduler


def _detect_input_keys(operator: Operator) -> Sequence[str]:
    # An instance of a family implies its base module is already loaded: look the bases up instead of importing
    # them (a first call used to pull in every attack and pre-aggregator module -- 17 ms in front of a 0.2 ms run).
    import sys

    pkg = __name__.rsplit(".", 3)[0]

    def _base(mod: str, name: str):
        return getattr(sys.modules.get(f"{pkg}.{mod}.base"), name, None)

    families = tuple(c for c in (_base("aggregators", "Aggregator"), _base("pre_aggregators", "PreAggregator")) if c)
    if families and isinstance(operator, families):
        return (operator.input_key,)
    attack = _base("attacks", "Attack")
    if attack is not None and isinstance(operator, attack):
        raise ValueError(
            f"Cannot auto-detect input keys for Attack {type(operator).__name__}. "
            "Attacks have variable input requirements. Please specify input_keys explicitly.")
    raise ValueError(f"Cannot auto-detect input keys for operator {type(operator).__name__}. "
                     "Please specify input_keys explicitly.")


class _RenamedInput(Operator):
    """Presents ``inputs[custom_key]`` to the wrapped operator under its own ``input_key``."""

    def __init__(self, wrapped: Operator, custom_key: str, op_key: str) -> None:
        self.wrapped_op = wrapped
        self.custom_key, self.op_key = custom_key, op_key
        self.name = wrapped.name
        self.supports_subtasks = wrapped.supports_subtasks
        self.supports_barriered_subtasks = wrapped.supports_barriered_subtasks
        self.max_subtasks_inflight = wrapped.max_subtasks_inflight

    def _remap(self, inputs: Mapping[str, Any]) -> dict:
        if self.custom_key not in inputs:
            raise KeyError(f"Missing input key {self.custom_key!r}")
        return {self.op_key: inputs[self.custom_key]}

    def compute(self, inputs, *, context: OpContext):
        return self.wrapped_op.compute(self._remap(inputs), context=context)

    async def run(self, inputs, *, context: OpContext, pool):
        # remap once so the wrapped operator sees ONE inputs mapping for the whole invocation
        return await self.wrapped_op.run(self._remap(inputs), context=context, pool=pool)


class OperatorExecutor:
    """Reusable single-operator runner: ``async with OperatorExecutor(op, pool_config=...) as ex`` keeps
    the actor pool, the one-node graph and its scheduler alive across ``await ex.run(inputs)`` calls
    (reference engine/graph/executor.py:70-294)."""

    def __init__(self, operator: Operator, *, input_keys: Optional[Sequence[str]] = None,
                 pool_config: Union[ActorPoolConfig, Sequence[ActorPoolConfig], None] = None,
                 node_name: Optional[str] = None):
        if not isinstance(operator, Operator):
            raise TypeError(f"operator must be an Operator instance, got {type(operator)}")
        self.operator = operator
        self.pool_config = pool_config
        self.node_name = node_name or operator.name
        self.input_keys = tuple(input_keys) if input_keys is not None else tuple(_detect_input_keys(operator))
        self._operator_input_key = getattr(operator, "input_key", None) if len(self.input_keys) == 1 else None
        self._needs_input_mapping = (self._operator_input_key is not None
                                     and self.input_keys[0] != self._operator_input_key)
        self._pool: Optional[ActorPool] = None
        self._pool_managed = False
        self._graph = None
        self._scheduler: Optional[NodeScheduler] = None

    async def __aenter__(self):
        if self.pool_config is not None:
            await self._ensure_pool()
        return self

    async def __aexit__(self, exc_type, exc_val, exc_tb):
        await self._cleanup_pool()

    async def _ensure_pool(self) -> None:
        if self._pool is not None:
            return
        cfgs = [self.pool_config] if isinstance(self.pool_config, ActorPoolConfig) else list(self.pool_config)
        self._pool = ActorPool(cfgs)
        await self._pool.start()
        self._pool_managed = True

    async def _cleanup_pool(self) -> None:
        if self._pool is not None and self._pool_managed:
            await self._pool.shutdown()
            self._pool = None
            self._pool_managed = False
            self._scheduler = None

    def _make_scheduler(self) -> NodeScheduler:
        op = self.operator
        if self._needs_input_mapping:
            op = _RenamedInput(op, self.input_keys[0], self._operator_input_key)
        graph = make_single_operator_graph(node_name=self.node_name, operator=op,
                                           input_keys=self.input_keys)
        self._graph = graph
        return NodeScheduler(graph, pool=self._pool)

    async def run(self, inputs: Mapping[str, Any]) -> Any:
        if self.pool_config is not None:
            await self._ensure_pool()
        reuse = self.pool_config is not None
        sched = self._scheduler if (reuse and self._scheduler is not None
                                    and self._scheduler.pool is self._pool) else None
        if sched is None:
            sched = self._make_scheduler()
            if reuse:
                self._scheduler = sched
        return (await sched.run(inputs))[self.node_name]


async def run_operator(operator: Operator, inputs: Mapping[str, Any], *,
                       pool_config: Union[ActorPoolConfig, Sequence[ActorPoolConfig], None] = None,
                       input_keys: Optional[Sequence[str]] = None) -> Any:
    """Run one operator once and return its result.

    ``operator``: an :class:`Aggregator` / :class:`PreAggregator` (its input key -- ``"gradients"`` /
    ``"vectors"`` -- is detected) or any :class:`Operator` together with ``input_keys``; attacks are refused
    (they take several differently-routed inputs: wire them into a graph).  ``inputs``: mapping from input
    key to value, e.g. ``{"gradients": [tensor, ...]}``.  ``pool_config``: one ``ActorPoolConfig`` or a
    sequence of them; an ``ActorPool`` is started for this call, the operator is decomposed into subtasks
    over it, and the pool is shut down again (use :class:`OperatorExecutor` to keep the pool across calls).
    Without ``pool_config`` the operator computes directly in this process.

    Returns whatever the operator returns (a tensor for aggregators, a list of tensors for
    pre-aggregators) on the input's device.
    """
    if pool_config is None:
        # no pool to start or stop: the one-node graph is kept per (operator, keys), so a
        # repeated call costs what ``scheduler.run`` on a hand-built graph costs (the pattern this helper replaces)
        try:
            per_op = _DIRECT.setdefault(operator, {})
        except TypeError:                      # not weak-referenceable / unhashable: no cache
            per_op = {}
        key = tuple(input_keys) if input_keys is not None else None
        cached = per_op.get(key)
        if cached is None:
            ex = OperatorExecutor(operator, input_keys=input_keys)
            cached = per_op[key] = (ex._make_scheduler().graph, ex.node_name)
        graph, name = cached
        return (await NodeScheduler(graph, pool=None).run(inputs))[name]
    async with OperatorExecutor(operator, pool_config=pool_config, input_keys=input_keys) as ex:
        return await ex.run(inputs)


_DIRECT: "weakref.WeakKeyDictionary[Operator, dict]" = weakref.WeakKeyDictionary()


__all__ = ["OperatorExecutor", "run_operator"]
