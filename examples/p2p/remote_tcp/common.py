"""Shared pieces of the remote-TCP P2P examples: node-list parsing and the per-node gossip loop."""
from __future__ import annotations

import asyncio
import json
import os
import sys
from typing import Dict, List

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from byzpy_b200.engine.graph.pool import ActorPoolConfig  # noqa: E402
from byzpy_b200.engine.node.application import ByzantineNodeApplication, HonestNodeApplication  # noqa: E402
from byzpy_b200.engine.node.decentralized import DecentralizedNode  # noqa: E402
from byzpy_b200.engine.peer_to_peer.topology import Topology  # noqa: E402


def load_config(path: str) -> dict:
    with open(path) as f:
        if path.endswith((".yaml", ".yml")):
            import yaml

            return yaml.safe_load(f)
        return json.load(f)


def add_training_flags(ap) -> None:
    """The per-node flags of the reference's clients (examples/p2p/remote_tcp/client.py:473-521,
    mesh_client.py): they override what the node list says."""
    ap.add_argument("--node-type", choices=["honest", "byzantine"], default=None,
                    help="this node's role (default: the node list's `role`, else honest)")
    ap.add_argument("--rounds", type=int, default=None)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--lr", type=float, default=0.05)
    ap.add_argument("--data-root", default="./data")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--byz-scale", type=float, default=-1.0)


def apply_training_flags(cfg: dict, a, node_id: str) -> dict:
    cfg = dict(cfg)
    if a.rounds is not None:
        cfg["rounds"] = a.rounds
    if a.node_type is not None:
        cfg["nodes"] = [dict(e, role=a.node_type) if str(e["id"]) == node_id else e for e in cfg["nodes"]]
    cfg["train"] = {"batch_size": a.batch_size, "lr": a.lr, "data_root": a.data_root, "seed": a.seed,
                    "byz_scale": a.byz_scale, "data_shard": getattr(a, "data_shard", None)}
    return cfg


def topology_of(cfg: dict) -> Topology:
    n = len(cfg["nodes"])
    return Topology.complete(n) if cfg.get("topology", "complete") == "complete" else Topology.ring(n, 1)


def id_map(cfg: dict) -> Dict[int, str]:
    return {i: str(e["id"]) for i, e in enumerate(cfg["nodes"])}


def make_node(cfg: dict, node_id: str, context) -> DecentralizedNode:
    entry = next(e for e in cfg["nodes"] if str(e["id"]) == node_id)
    cls = HonestNodeApplication if entry.get("role", "honest") == "honest" else ByzantineNodeApplication
    app = cls(name=f"node{node_id}", actor_pool=[ActorPoolConfig(backend="thread", count=1)])
    return DecentralizedNode(node_id=node_id, application=app, context=context, topology=topology_of(cfg),
                             node_id_map=id_map(cfg))


async def gossip(node: DecentralizedNode, cfg: dict, role: str, *, settle: float = 1.0,
                 recv_timeout: float = 15.0) -> None:
    """The node's own training loop (honest) or attack loop (byzantine); returns after cfg['rounds']."""
    from examples.p2p.nodes import P2PByzNode, P2PHonestNode
    from byzpy_b200.utils.data import shard_indices

    ids: List[str] = [str(e["id"]) for e in cfg["nodes"]]
    honest_ids = [str(e["id"]) for e in cfg["nodes"] if e.get("role", "honest") == "honest"]
    rounds = int(cfg.get("rounds", 5))
    train = cfg.get("train") or {}
    lr = float(train.get("lr", 0.05))
    inbox, arrived = [], asyncio.Event()

    async def on_model(frm, payload):
        inbox.append((frm, payload["vector"]))
        arrived.set()

    node.register_message_handler("model", on_model)
    await asyncio.sleep(settle)                 # let the other processes come up / connect

    async def collect(expected: int):
        loop = asyncio.get_running_loop()
        deadline = loop.time() + recv_timeout
        while len(inbox) < expected and loop.time() < deadline:
            arrived.clear()
            try:
                await asyncio.wait_for(arrived.wait(), timeout=max(0.01, deadline - loop.time()))
            except asyncio.TimeoutError:
                break
        got = [v for _, v in inbox]
        inbox.clear()
        return got

    if role == "honest":
        me = honest_ids.index(node.node_id)
        shard = me if train.get("data_shard") is None else int(train["data_shard"]) % len(honest_ids)
        worker = P2PHonestNode(indices=shard_indices(6000, len(honest_ids))[shard], seed=int(train.get("seed", 0)) + me,
                               batch_size=int(train.get("batch_size", 64)), data_root=train.get("data_root", "./data"))
        n_in = len(node.get_in_neighbors())
        for r in range(1, rounds + 1):
            own = worker.p2p_half_step(lr)
            await node.broadcast_message("model", {"vector": own})
            received = await collect(n_in)
            if len(received) + 1 > 2 * worker.p2p_agg.f:
                worker.p2p_aggregate_and_set(own, received)
            print(f"[node {node.node_id}] round {r}: {len(received)} neighbour vectors, "
                  f"|theta| = {worker.get_param_vector().norm().item():.4f}", flush=True)
    else:
        attacker = P2PByzNode(scale=float(train.get("byz_scale", -1.0)))
        n_h_in = len([i for i in node.get_in_neighbors() if str(i) in honest_ids])
        for r in range(1, rounds + 1):
            seen = await collect(n_h_in)
            if seen:
                out = attacker.p2p_broadcast_vector(neighbor_vectors=seen, like=seen[0])
                await node.broadcast_message("model", {"vector": out})
            print(f"[node {node.node_id}] round {r}: attacked with {len(seen)} honest vectors", flush=True)
    _ = ids, torch
