"""Generic operators wrapping plain callables (reference engine/graph/ops.py:10-99)."""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterable, Mapping, Sequence

from .graph import ComputationGraph, GraphNode, graph_input
from .operator import OpContext, Operator
from .subtask import SubTask


def _bind(mapping: Mapping[str, str], inputs: Mapping[str, Any], who: str) -> Dict[str, Any]:
    bound: Dict[str, Any] = {}
    for param, source in mapping.items():
        if source not in inputs:
            raise KeyError(f"{who} missing required input {source!r} for parameter {param!r}")
        bound[param] = inputs[source]
    return bound


class CallableOp(Operator):
    """Calls ``fn(**{param: inputs[source]})`` in the scheduler's process."""

    name = "callable"

    def __init__(self, fn: Callable[..., Any], *, input_mapping: Mapping[str, str]) -> None:
        self.fn = fn
        self.input_mapping = dict(input_mapping)

    def compute(self, inputs: Mapping[str, Any], *, context: OpContext) -> Any:
        return self.fn(**_bind(self.input_mapping, inputs, "CallableOp"))


def _invoke_remote_callable(fn: Callable[..., Any], kwargs: Mapping[str, Any]) -> Any:
    return fn(**kwargs)


class RemoteCallableOp(Operator):
    """Ships the call to a pool worker as a single subtask (falls back to local without a pool)."""

    name = "remote_callable"
    supports_subtasks = True

    def __init__(self, fn: Callable[..., Any], *, input_mapping: Mapping[str, str]) -> None:
        self.fn = fn
        self.input_mapping = dict(input_mapping)

    def compute(self, inputs: Mapping[str, Any], *, context: OpContext) -> Any:
        return self.fn(**_bind(self.input_mapping, inputs, "RemoteCallableOp"))

    def create_subtasks(self, inputs: Mapping[str, Any], *, context: OpContext) -> Iterable[SubTask]:
        kwargs = _bind(self.input_mapping, inputs, "RemoteCallableOp")
        return [SubTask(fn=_invoke_remote_callable, args=(self.fn, kwargs), kwargs={})]

    def reduce_subtasks(self, partials: Sequence[Any], inputs: Mapping[str, Any], *,
                        context: OpContext) -> Any:
        if not partials:
            raise RuntimeError("RemoteCallableOp expected exactly one partial result.")
        return partials[0]


def make_single_operator_graph(*, node_name: str, operator: Operator,
                               input_keys: Sequence[str]) -> ComputationGraph:
    """The one-node graph ``operator(input_keys...) -> node_name``: every input key becomes a run-time graph input."""
    node = GraphNode(name=node_name, op=operator, inputs={k: graph_input(k) for k in input_keys})
    return ComputationGraph([node], outputs=[node_name])


__all__ = ["CallableOp", "RemoteCallableOp", "make_single_operator_graph"]
