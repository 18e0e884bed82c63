"""Engine package (reference engine/__init__.py:1-7)."""
from .graph.ops import CallableOp, RemoteCallableOp, make_single_operator_graph
from .node_cluster import NodeCluster
from .node_runner import NodeRunner

__all__ = ["CallableOp", "RemoteCallableOp", "make_single_operator_graph", "NodeCluster", "NodeRunner"]
