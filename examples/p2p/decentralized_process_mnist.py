"""Decentralized peer-to-peer training with ``DecentralizedPeerToPeer``: every node is a
``DecentralizedNode`` (own scheduler + message router), optionally in its own OS process
(``ProcessContext``), gossiping half-step models over a ring / complete topology.

    python examples/p2p/decentralized_process_mnist.py [--rounds 20] [--context inprocess|process]
"""
from __future__ import annotations

import argparse
import asyncio
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from examples.p2p.nodes import P2PByzNode, P2PHonestNode  # noqa: E402

from byzpy_b200.engine.node.actors import ByzantineNodeActor, HonestNodeActor  # noqa: E402
from byzpy_b200.engine.node.context import InProcessContext, ProcessContext  # noqa: E402
from byzpy_b200.engine.peer_to_peer.runner import DecentralizedPeerToPeer  # noqa: E402
from byzpy_b200.engine.peer_to_peer.topology import Topology  # noqa: E402
from byzpy_b200.models import SmallCNN  # noqa: E402
from byzpy_b200.utils.data import evaluate, mnist_like, shard_indices  # noqa: E402


async def main(rounds: int, context: str, topology: str):
    n_h, n_b = 4, 1
    shards = shard_indices(6000, n_h)
    hon = [await HonestNodeActor.spawn(P2PHonestNode, backend="thread", kwargs=dict(indices=shards[i], seed=i))
           for i in range(n_h)]
    byz = [await ByzantineNodeActor.spawn(P2PByzNode, backend="thread") for _ in range(n_b)]
    topo = Topology.complete(n_h + n_b) if topology == "complete" else Topology.ring(n_h + n_b, 2)
    factory = (lambda node_id, idx: ProcessContext()) if context == "process" else (
        lambda node_id, idx: InProcessContext())
    p2p = DecentralizedPeerToPeer(hon, byz, topo, lr=0.05, context_factory=factory, recv_timeout=30.0)
    await p2p.start()
    xt, yt = mnist_like(2000, train=False)
    probe = SmallCNN()
    for r in range(1, rounds + 1):
        await p2p.run_round_async()
        if r % max(1, rounds // 5) == 0:
            probe.load_state_dict(await hon[0].dump_state_dict(), strict=True)
            loss, acc = evaluate(probe, xt, yt, torch.device("cpu"))
            print(f"[round {r:04d}] node0 test loss={loss:.4f} acc={acc:.4f}")
    await p2p.stop()
    for a in hon + byz:
        await a.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--context", default="inprocess", choices=["inprocess", "process"])
    ap.add_argument("--topology", default="complete")
    a = ap.parse_args()
    asyncio.run(main(a.rounds, a.context, a.topology))
