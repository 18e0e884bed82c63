"""``PeerToPeer``: the user-facing P2P training facade
(reference engine/peer_to_peer/train.py:17-86)."""
from __future__ import annotations

from typing import Any, Callable, List, Optional

from ..node.context import NodeContext
from .runner import DecentralizedPeerToPeer
from .topology import Topology


class PeerToPeer:
    def __init__(self, honest_nodes: List[Any], byzantine_nodes: Optional[List[Any]], topology: Topology, *,
                 lr: float = 0.05, channel_name: str = "p2p",
                 context_factory: Optional[Callable[[str, int], NodeContext]] = None):
        self.channel_name = channel_name  # kept for API compatibility
        self._runner = DecentralizedPeerToPeer(honest_nodes, byzantine_nodes or [], topology, lr=lr,
                                               context_factory=context_factory)

    @property
    def runner(self) -> DecentralizedPeerToPeer:
        return self._runner

    async def bootstrap(self) -> None:
        await self._runner.start()

    async def round(self) -> None:
        await self._runner.run_round_async()

    async def shutdown(self) -> None:
        await self._runner.stop()


__all__ = ["PeerToPeer"]
