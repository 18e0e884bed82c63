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
    """Fluent construction of a :class:`~byzpy_b200.engine.graph.graph.ComputationGraph`.

    ``input(name)`` declares run-time data and returns a :class:`LazyNode`; ``LazyNode.apply(operator)`` records a node
    fed by it and returns the next lazy node; ``build(outputs)`` validates and returns the graph.  Nothing executes
    until the graph is given to a scheduler.

    Examples
    --------
    >>> from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
    >>> from byzpy_b200.pre_aggregators import Clipping
    >>> from byzpy_b200.engine.graph.lazy import GraphBuilder
    >>> b = GraphBuilder()
    >>> out = b.input("vectors").apply(Clipping(threshold=1.0)).apply(CoordinateWiseMedian(), name="agg")
    >>> g = b.build(outputs=[out.key])
    >>> [n.name for n in g.nodes_in_order()]
    ['pre-agg/clipping_0', 'agg']
    """

    def __init__(self) -> None:
        self._nodes: Dict[str, GraphNode] = {}
        self._inputs: Dict[str, GraphInput] = {}
        self._node_counter = 0

    def input(self, name: str) -> "LazyNode":
        """Declare run-time data called ``name`` and return its lazy handle."""
        self._inputs.setdefault(name, GraphInput(name))
        return LazyNode(builder=self, key=name, is_input=True)

    def build(self, outputs: Sequence[str]) -> ComputationGraph:
        """Validate what was recorded and return the graph whose results are the nodes named in ``outputs``."""
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
    """Handle on a not-yet-computed value inside a :class:`GraphBuilder` (an input or a recorded node).

    ``apply(operator, input_key=None, extra_inputs=None, name=None)`` wires this value into ``operator`` under
    ``input_key`` (default: the operator's own ``input_key``, e.g. ``"gradients"`` for aggregators and ``"vectors"`` for
    pre-aggregators), optionally with further inputs (lazy nodes or names), and returns the handle of the new node.
    ``key`` is the node (or input) name.
    """

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
        """Record ``operator`` fed by this value (under ``input_key``, default the operator's own) plus ``extra_inputs``
        (lazy nodes or node names); returns the handle of the new node, named ``name`` or ``"<operator.name>_<k>"``.
        """
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
