"""``DecentralizedCluster``: add / start / shut down a set of decentralized nodes
(reference engine/node/cluster.py:12-111).  The default context is ``ProcessContext``; the
``index -> node_id`` map (insertion order) is what lets integer topologies route to string ids."""
from __future__ import annotations

from typing import Any, Dict, Mapping, Optional, Union

from .application import NodeApplication
from .context import NodeContext, ProcessContext
from .decentralized import DecentralizedNode

NodeId = Union[int, str]


class DecentralizedCluster:
    """A set of :class:`~byzpy_b200.engine.node.decentralized.DecentralizedNode` objects managed together.

    ``await add_node(node_id, application, topology=None, context=None, metadata=None)`` creates a node (default context:
    a :class:`~byzpy_b200.engine.node.context.ProcessContext`) and returns it; the order of insertion defines the
    topology index of each id.  ``await start_all()``, ``await shutdown_all()``, ``get_node(id)``,
    ``await remove_node(id)``; ``nodes`` maps ids to nodes.
    """

    def __init__(self) -> None:
        self.nodes: Dict[NodeId, DecentralizedNode] = {}
        self._node_id_map: Dict[int, NodeId] = {}

    async def add_node(self, node_id: NodeId, application: NodeApplication, topology: Any = None,
                       context: Optional[NodeContext] = None,
                       metadata: Optional[Mapping[str, Any]] = None) -> DecentralizedNode:
        """Create a node (context default: a fresh :class:`ProcessContext`) and register it; returns the node.
        ``ValueError`` when the id exists.
        """
        if node_id in self.nodes:
            raise ValueError(f"Node {node_id!r} already exists in cluster")
        if context is None:
            context = ProcessContext()
        index = len(self.nodes)
        self._node_id_map[index] = node_id
        node = DecentralizedNode(node_id=node_id, application=application, context=context,
                                 topology=topology, metadata=metadata,
                                 node_id_map=dict(self._node_id_map))
        self.nodes[node_id] = node
        self._update_node_id_maps()
        return node

    def _update_node_id_maps(self) -> None:
        snapshot = dict(self._node_id_map)
        for node in self.nodes.values():
            node._node_id_map = snapshot
            r = node.message_router
            r._node_id_map = dict(snapshot)
            r._reverse_id_map = {v: k for k, v in snapshot.items()}

    async def start_all(self) -> None:
        """Start every node, in insertion order."""
        self._update_node_id_maps()
        for node in self.nodes.values():
            await node.start()

    async def shutdown_all(self) -> None:
        """Shut every node down and forget them."""
        for node in list(self.nodes.values()):
            await node.shutdown()
        self.nodes.clear()
        self._node_id_map.clear()

    def get_node(self, node_id: NodeId) -> Optional[DecentralizedNode]:
        """The node with this id, or ``None``."""
        return self.nodes.get(node_id)

    async def remove_node(self, node_id: NodeId) -> None:
        """Shut one node down and remove it (unknown ids are ignored); the remaining nodes are re-indexed."""
        node = self.nodes.pop(node_id, None)
        if node is None:
            return
        await node.shutdown()
        self._node_id_map = {i: nid for i, nid in enumerate(self.nodes)}
        self._update_node_id_maps()


__all__ = ["DecentralizedCluster"]
