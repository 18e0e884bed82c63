"""Computation graph: a validated DAG of operator nodes.

Contract of the reference structure (reference engine/graph/graph.py:23-131): nodes have unique
names; an input mapping value is a ``GraphInput`` (external data), a node name (edge), or a
``MessageSource`` (resolved by a message-aware scheduler; neither an input nor an edge).
Duplicate names, unknown dependencies, unknown outputs and cycles raise ``ValueError``.
``outputs`` defaults to the last node of the topological order.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Mapping, Optional, Sequence, Set

from .operator import Operator


@dataclass(frozen=True)
class GraphInput:
    """Opaque reference to data supplied by the application when the graph runs."""

    name: str

    @classmethod
    def from_message(cls, message_type: str, field: Optional[str] = None,
                     timeout: Optional[float] = None):
        from .scheduler import MessageSource

        return MessageSource(message_type=message_type, field=field, timeout=timeout)


def graph_input(name: str) -> GraphInput:
    """Placeholder for data handed to the scheduler at run time under ``name``."""
    return GraphInput(name)


@dataclass(frozen=True)
class GraphNode:
    """One node of a :class:`ComputationGraph`: an operator and where each of its inputs comes from.

    Parameters
    ----------
    name : str
        Unique name; other nodes refer to this node's result by it, schedulers return results keyed by it.
    op : Operator
        The operator to run.
    inputs : mapping of str to source
        For each input key of the operator: a :func:`graph_input` (data given at run time), the name of another node
        (an edge), or a ``MessageSource`` (a message a :class:`MessageAwareNodeScheduler` waits for).
    """

    name: str
    op: Operator
    inputs: Mapping[str, Any] = field(default_factory=dict)


def _is_message_source(dep: Any) -> bool:
    return hasattr(dep, "message_type")


class ComputationGraph:
    """A validated directed acyclic graph of operator nodes.

    Parameters
    ----------
    nodes : sequence of GraphNode
        At least one; names must be unique, every edge must name an existing node, cycles are rejected
        (``ValueError``).
    outputs : sequence of str, optional
        Names of the nodes whose results a scheduler returns; default: the last node in topological order.

    Attributes
    ----------
    required_inputs : names of the :func:`graph_input` placeholders the run must be fed.
    outputs : the output node names.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
    >>> from byzpy_b200.pre_aggregators import Clipping
    >>> from byzpy_b200.engine.graph.graph import ComputationGraph, GraphNode, graph_input
    >>> g = ComputationGraph([
    ...     GraphNode("clip", Clipping(threshold=1.0), {"vectors": graph_input("grads")}),
    ...     GraphNode("agg", CoordinateWiseMedian(), {"gradients": "clip"}),
    ... ])
    >>> sorted(g.required_inputs), list(g.outputs)
    (['grads'], ['agg'])
    """

    def __init__(self, nodes: Sequence[GraphNode], *, outputs: Optional[Sequence[str]] = None) -> None:
        if not nodes:
            raise ValueError("ComputationGraph requires at least one node.")
        self._nodes: Dict[str, GraphNode] = {}
        for node in nodes:
            if node.name in self._nodes:
                raise ValueError(f"Duplicate graph node: {node.name}")
            self._nodes[node.name] = node
        self._edges: Dict[str, List[str]] = {}     # node -> names of the nodes it depends on
        external: Set[str] = set()
        for node in nodes:
            deps: List[str] = []
            for dep in node.inputs.values():
                if isinstance(dep, GraphInput):
                    external.add(dep.name)
                elif _is_message_source(dep):
                    continue
                elif dep in self._nodes:
                    if dep not in deps:
                        deps.append(dep)
                else:
                    raise ValueError(f"Node {node.name} depends on unknown node {dep!r}: the graph contains a cycle "
                                     f"or an unresolved dependency")
            self._edges[node.name] = deps
        self._order = self._toposort([n.name for n in nodes])
        self.outputs: List[str] = list(outputs) if outputs is not None else [self._order[-1]]
        for out in self.outputs:
            if out not in self._nodes:
                raise ValueError(f"Unknown output node: {out}")
        self.required_inputs = frozenset(external)

    # Kahn's algorithm, stable w.r.t. declaration order
    def _toposort(self, declared: List[str]) -> List[str]:
        pending = {name: len(self._edges[name]) for name in declared}
        children: Dict[str, List[str]] = {name: [] for name in declared}
        for name in declared:
            for dep in self._edges[name]:
                children[dep].append(name)
        frontier = [name for name in declared if pending[name] == 0]
        order: List[str] = []
        while frontier:
            name = frontier.pop(0)
            order.append(name)
            for child in children[name]:
                pending[child] -= 1
                if pending[child] == 0:
                    frontier.append(child)
        if len(order) != len(declared):
            raise ValueError("ComputationGraph contains a cycle; cannot determine order.")
        return order

    def nodes_in_order(self) -> Iterable[GraphNode]:
        """The nodes in a topological order (dependencies first)."""
        for name in self._order:
            yield self._nodes[name]

    def node(self, name: str) -> GraphNode:
        return self._nodes[name]

    def dependencies(self, name: str) -> List[str]:
        """Names of the nodes ``name`` takes inputs from."""
        return list(self._edges[name])

    def __len__(self) -> int:
        return len(self._nodes)

    def __contains__(self, name: str) -> bool:
        return name in self._nodes


__all__ = ["ComputationGraph", "GraphInput", "GraphNode", "graph_input"]
