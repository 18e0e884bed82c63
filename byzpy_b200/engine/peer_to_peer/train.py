"""``PeerToPeer``: the user-facing P2P training facade
(reference engine/peer_to_peer/train.py:17-86)."""
from __future__ import annotations

from typing import Any, Callable, List, Optional

from ..node.context import NodeContext
from .runner import DecentralizedPeerToPeer
from .topology import Topology


class PeerToPeer:
    """Message-driven P2P training, or -- when every node is a device node on CUDA -- the fused
    device round of :mod:`byzpy_b200.parallel.device_p2p` (multi-GPU: one process per GPU, each
    passes its LOCAL nodes plus ``layout=PeerLayout(n_honest, n_byz, world)``)."""

    def __init__(self, honest_nodes: List[Any], byzantine_nodes: Optional[List[Any]], topology: Topology, *,
                 lr: float = 0.05, channel_name: str = "p2p",
                 context_factory: Optional[Callable[[str, int], NodeContext]] = None, layout=None,
                 process_group=None, amp_dtype=None, use_cuda_graph: bool = True,
                 fused: Optional[bool] = None):
        self.channel_name = channel_name  # kept for API compatibility
        self.device_round = None
        self._hon, self._byz = list(honest_nodes), list(byzantine_nodes or [])
        if fused is not False and context_factory is None:
            self.device_round = self._try_device_round(topology, lr, layout, process_group, amp_dtype,
                                                       use_cuda_graph)
        if fused and self.device_round is None:
            raise RuntimeError("fused=True requested but the configuration has no fused device path")
        self._runner = None if self.device_round is not None else DecentralizedPeerToPeer(
            honest_nodes, byzantine_nodes or [], topology, lr=lr, context_factory=context_factory)

    def _try_device_round(self, topology, lr, layout, group, amp_dtype, use_cuda_graph):
        import torch

        from ..node.device import DeviceP2PByzantineNode, DeviceP2PHonestNode

        nodes = self._hon + self._byz
        if not nodes or not torch.cuda.is_available():
            return None
        if not all(isinstance(n, DeviceP2PHonestNode) for n in self._hon):
            return None
        if not all(isinstance(n, DeviceP2PByzantineNode) for n in self._byz):
            return None
        if any(n.device.type != "cuda" for n in nodes) or any(n.p2p_pre is not None for n in self._hon):
            return None
        from ...parallel.device_p2p import DeviceP2PRound, DevicePeer, PeerLayout

        if layout is None:
            layout = PeerLayout(len(self._hon), len(self._byz), 1)
        rank = 0
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            rank = dist.get_rank(group)
        gids = layout.local_ids(rank)
        local_nodes = self._hon + self._byz
        if len(gids) != len(local_nodes):
            raise ValueError("pass exactly this rank's nodes (honest first, then Byzantine)")
        peers = []
        for g, node in zip(gids, local_nodes):
            n_rows = len(dict.fromkeys(topology.in_.get(g, [])))
            if g < layout.n_honest:
                plan = node.p2p_agg.fused_plan(n_rows + 1)
                if plan is None:
                    return None
                peers.append(DevicePeer(role="honest", model=node.model, loss_fn=node.criterion, plan=plan,
                                        preprocess=node.preprocess, data=node.data, name=node.name))
            else:
                honest_in = len([j for j in dict.fromkeys(topology.in_.get(g, [])) if j < layout.n_honest])
                fold = node.attack.fold(honest_in)
                if fold is None or fold.kind not in ("virtual", "alias"):
                    return None
                peers.append(DevicePeer(role="byzantine", fold=fold, name=node.name))
        return DeviceP2PRound(peers, layout, topology, lr=lr, device=local_nodes[0].device, group=group,
                              amp_dtype=amp_dtype, use_cuda_graph=use_cuda_graph)

    @property
    def runner(self) -> DecentralizedPeerToPeer:
        """The message-driven runner behind the generic path (``None`` on the fused device path)."""
        return self._runner

    async def bootstrap(self) -> None:
        """Create and start the decentralized nodes (generic path); a no-op on the device path.  Call once before :meth:`round`."""
        if self._runner is not None:
            await self._runner.start()

    def step(self, batches=None):
        """Device path: one fused gossip round (asynchronous on the current CUDA stream)."""
        if self.device_round is None:
            raise RuntimeError("step() is only available on the fused device path; use round()")
        return self.device_round.step(batches)

    async def round(self) -> None:
        """One gossip round: every honest node takes a local half step, exchanges models with its neighbours and robustly
        aggregates what it received.
        """
        if self.device_round is not None:
            self.device_round.step()
            return
        await self._runner.run_round_async()

    async def shutdown(self) -> None:
        """Stop the nodes / release the device round."""
        if self.device_round is not None:
            self.device_round.close()
            self.device_round = None
        if self._runner is not None:
            await self._runner.stop()


__all__ = ["PeerToPeer"]
