"""Node runtime (reference package ``byzpy.engine.node``, exports of node/__init__.py:17-38)."""
from ..graph.ops import CallableOp, RemoteCallableOp, make_single_operator_graph
from .application import (ByzantineNodeApplication, HonestNodeApplication, NodeApplication,
                          NodePipeline)
from .cluster import DecentralizedCluster
from .context import InProcessContext, MeshRemoteContext, NodeContext, ProcessContext, RemoteContext
from .decentralized import DecentralizedNode
from .distributed import DistributedByzantineNode, DistributedHonestNode
from .remote_client import RemoteNodeClient, deserialize_message, serialize_message
from .remote_server import RemoteNodeServer
from .router import MessageRouter

__all__ = [
    "NodeApplication", "NodePipeline", "HonestNodeApplication", "ByzantineNodeApplication",
    "CallableOp", "RemoteCallableOp", "make_single_operator_graph", "NodeContext",
    "InProcessContext", "ProcessContext", "RemoteContext", "MeshRemoteContext", "DecentralizedNode",
    "DistributedHonestNode", "DistributedByzantineNode", "DecentralizedCluster", "MessageRouter",
    "RemoteNodeServer", "RemoteNodeClient", "serialize_message", "deserialize_message",
]
