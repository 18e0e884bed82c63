"""Topology-aware message routing for one node (reference engine/node/router.py:10-260).

Node ids may be topology integers or strings mapped through ``node_id_map`` (topology index ->
string id).  ``route_direct`` rejects self-sends and non-neighbours, ``route_broadcast`` sends to
the de-duplicated out-neighbours and tolerates individual failures, ``route_multicast`` validates
every target first, ``route_reply`` answers the sender of a message.  ``topology=None`` allows
everything (and has no neighbours to broadcast to).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Union

NodeId = Union[int, str]


class MessageRouter:
    """Sends a node's outgoing messages through its context, as far as the topology allows.

    All ``route_*`` methods are coroutines taking the message type, the payload and the
    :class:`~byzpy_b200.engine.node.context.NodeContext` to send through: ``route_direct(target, ...)`` refuses
    self-sends and targets that are not out-neighbours (``ValueError``), ``route_broadcast(...)`` sends to the
    de-duplicated out-neighbours and carries on past individual failures, ``route_multicast(targets, ...)`` validates
    every target before sending, ``route_reply(original_message, ...)`` answers its sender.  ``get_out_neighbors`` /
    ``get_in_neighbors`` / ``can_send_to`` answer questions about the neighbourhood.  Ids can be topology integers
    or strings translated through ``node_id_map``; without a topology everything is allowed and a broadcast reaches
    nobody.

    Examples
    --------
    >>> from byzpy_b200.engine.node.router import MessageRouter
    >>> from byzpy_b200.engine.peer_to_peer.topology import Topology
    >>> r = MessageRouter(topology=Topology.ring(4, 1), node_id="n0", node_id_map={i: f"n{i}" for i in range(4)})
    >>> r.get_out_neighbors(), r.can_send_to("n2")
    (['n1', 'n3'], False)
    """

    def __init__(self, *, topology: Optional[Any] = None, node_id: NodeId,
                 node_id_map: Optional[Dict[int, str]] = None):
        self.topology = topology
        self.node_id = node_id
        self._node_id_map: Dict[int, str] = dict(node_id_map or {})
        self._reverse_id_map: Dict[str, int] = {v: k for k, v in self._node_id_map.items()}
        if topology is not None and isinstance(node_id, int) and not (0 <= node_id < topology.n):
            raise ValueError(f"node_id {node_id} is not in topology (n={topology.n})")

    # ---- id translation ------------------------------------------------------------------
    def _to_internal_id(self, node_id: NodeId) -> int:
        if isinstance(node_id, int):
            return node_id
        return self._reverse_id_map.get(node_id, -1)

    def _to_external_id(self, internal_id: int) -> NodeId:
        return self._node_id_map.get(internal_id, internal_id)

    def _self_index(self) -> int:
        return self._to_internal_id(self.node_id)

    # ---- neighbourhood ---------------------------------------------------------------------
    def get_out_neighbors_internal(self) -> List[int]:
        if self.topology is None:
            return []
        return list(self.topology.out.get(self._self_index(), []))

    def get_out_neighbors(self) -> List[NodeId]:
        return [self._to_external_id(i) for i in self.get_out_neighbors_internal()]

    def get_in_neighbors(self) -> List[NodeId]:
        if self.topology is None:
            return []
        return [self._to_external_id(i) for i in self.topology.in_.get(self._self_index(), [])]

    def can_send_to(self, target_node_id: NodeId) -> bool:
        if self.topology is None:
            return True
        idx = self._to_internal_id(target_node_id)
        if idx < 0:
            return False
        return idx in self.get_out_neighbors_internal()

    # ---- routing ---------------------------------------------------------------------------
    async def route_direct(self, target_node_id: NodeId, message_type: str, payload: Any, context) -> None:
        if target_node_id == self.node_id:
            raise ValueError("cannot send to self")
        if self.topology is not None and not self.can_send_to(target_node_id):
            raise ValueError(f"Target {target_node_id} is not a neighbor of {self.node_id}")
        await context.send_message(target_node_id, message_type, payload)

    async def route_broadcast(self, message_type: str, payload: Any, context) -> None:
        if context is None:
            return
        for neighbour in dict.fromkeys(self.get_out_neighbors()):
            try:
                await context.send_message(neighbour, message_type, payload)
            except Exception:
                continue  # one dead neighbour must not stop the broadcast

    async def route_multicast(self, target_node_ids: List[NodeId], message_type: str, payload: Any,
                              context) -> None:
        if not target_node_ids:
            return
        if self.topology is not None:
            for t in target_node_ids:
                if not self.can_send_to(t):
                    raise ValueError(f"Target {t} is not a neighbor of {self.node_id}")
        if context is None:
            return
        for t in target_node_ids:
            await context.send_message(t, message_type, payload)

    async def route_reply(self, original_message: Dict[str, Any], message_type: str, payload: Any,
                          context) -> None:
        sender = original_message.get("from")
        if sender is None:
            raise ValueError("Original message has no 'from' field")
        await self.route_direct(sender, message_type, payload, context)

    async def route_message(self, target_node_id: NodeId, message_type: str, payload: Any, context) -> None:
        await self.route_direct(target_node_id, message_type, payload, context)


__all__ = ["MessageRouter"]
