"""Fluent lazy graph construction (reference engine/graph/lazy.py:24-229).

``builder.input(name)`` returns a :class:`LazyNode`; ``node.apply(op, ...)`` records a graph node
and returns a new lazy node.  Nothing executes until the built graph is handed to a scheduler.
Auto-generated node names are ``"{operator.name}_{counter}"``.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Union

from .graph import ComputationGraph, GraphInput, GraphNode
from .operator import Operator


class GraphBuilder:
    def __init__(self) -> None:
        self._nodes: Dict[str, GraphNode] = {}
        self._inputs: Dict[str, GraphInput] = {}
        self._node_counter = 0

    def input(self, name: str) -> "LazyNode":
        self._inputs.setdefault(name, GraphInput(name))
        return LazyNode(builder=self, key=name, is_input=True)

    def build(self, outputs: Sequence[str]) -> ComputationGraph:
        if not self._nodes:
            raise ValueError("GraphBuilder requires at least one node")
        for name in outputs:
            if name not in self._nodes:
                raise ValueError(f"Unknown output node: {name}")
        return ComputationGraph(list(self._nodes.values()), outputs=list(outputs))

    def _generate_node_name(self, operator: Operator) -> str:
        name = f"{operator.name}_{self._node_counter}"
        self._node_counter += 1
        return name

    def _ref(self, node: "LazyNode") -> Union[str, GraphInput]:
        return self._inputs[node.key] if node._is_input else node.key


class LazyNode:
    def __init__(self, builder: GraphBuilder, key: str, is_input: bool = False) -> None:
        self._builder = builder
        self._key = key
        self._is_input = is_input

    @property
    def key(self) -> str:
        return self._key

    def apply(self, operator: Operator, *, input_key: Optional[str] = None,
              extra_inputs: Optional[Dict[str, Union["LazyNode", str]]] = None,
              name: Optional[str] = None) -> "LazyNode":
        if not isinstance(operator, Operator):
            raise TypeError(f"operator must be an Operator instance, got {type(operator).__name__}")
        b = self._builder
        if input_key is None:
            input_key = getattr(operator, "input_key", "vectors")
        out = name if name is not None else b._generate_node_name(operator)
        wiring: Dict[str, Any] = {input_key: b._ref(self)}
        for arg, dep in (extra_inputs or {}).items():
            wiring[arg] = b._ref(dep) if isinstance(dep, LazyNode) else dep
        b._nodes[out] = GraphNode(name=out, op=operator, inputs=wiring)
        return LazyNode(builder=b, key=out)


__all__ = ["GraphBuilder", "LazyNode"]
