"""Node runtime: applications, decentralized nodes, message contexts, clusters (counterpart of the
reference package ``byzpy.engine.node``; same export list as its node/__init__.py:17-38, plus
``MeshRemoteContext``).  Resolved from a name -> module table."""
from importlib import import_module as _import_module

_WHERE = {
    # pipelines hosted by a node
    "NodeApplication": ".application", "NodePipeline": ".application",
    "HonestNodeApplication": ".application", "ByzantineNodeApplication": ".application",
    # single-operator graph helpers re-exported for convenience
    "CallableOp": "..graph.ops", "RemoteCallableOp": "..graph.ops", "make_single_operator_graph": "..graph.ops",
    # message transports behind a node
    "NodeContext": ".context", "InProcessContext": ".context", "ProcessContext": ".context",
    "RemoteContext": ".context", "MeshRemoteContext": ".context",
    # nodes and their containers
    "DecentralizedNode": ".decentralized", "DistributedHonestNode": ".distributed",
    "DistributedByzantineNode": ".distributed", "DecentralizedCluster": ".cluster", "MessageRouter": ".router",
    # hub-and-spoke networking
    "RemoteNodeServer": ".remote_server", "RemoteNodeClient": ".remote_client",
    "serialize_message": ".remote_client", "deserialize_message": ".remote_client",
}

for _name, _module in _WHERE.items():
    globals()[_name] = getattr(_import_module(_module, __name__), _name)

__all__ = list(_WHERE)
