"""Fully decentralised, autonomous peer-to-peer training (no coordinator drives the rounds).

Every node is a ``DecentralizedNode``.  Its model is built *inside the node* by an ``init_callback``
(so with ``--context process`` the model only ever exists in the child process), and an autonomous
task runs the node's own loop:

  honest   repeat: half step on the local shard -> broadcast theta^{t+1/2} to the out-neighbours ->
           wait for one vector per in-neighbour (or a timeout) -> robust aggregate -> write back
  byzantine repeat: wait for the honest in-neighbours' vectors -> Empire attack -> broadcast

The three stages are registered as node pipelines (``half_step`` / ``aggregate`` / ``update_model``)
and executed through the node's scheduler, like the reference's
examples/p2p/decentralized_autonomous_mnist.py.

    python examples/p2p/decentralized_autonomous_mnist.py [--rounds 10] [--context inprocess|process]
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from byzpy_b200.engine.graph.ops import CallableOp, make_single_operator_graph  # noqa: E402
from byzpy_b200.engine.graph.pool import ActorPoolConfig  # noqa: E402
from byzpy_b200.engine.node.application import ByzantineNodeApplication, HonestNodeApplication  # noqa: E402
from byzpy_b200.engine.node.cluster import DecentralizedCluster  # noqa: E402
from byzpy_b200.engine.node.context import InProcessContext, ProcessContext  # noqa: E402
from byzpy_b200.engine.peer_to_peer.topology import Topology  # noqa: E402

N_HONEST, N_BYZ = 4, 1


def _pipeline(name, fn, keys):
    return make_single_operator_graph(node_name=name, operator=CallableOp(fn, input_mapping={k: k for k in keys}),
                                      input_keys=tuple(keys))


def _go_event(node) -> asyncio.Event:
    """Start barrier: the launcher sends one "go" message once every node of the cluster runs."""
    ev = asyncio.Event()

    async def on_go(frm, payload):
        ev.set()

    node.register_message_handler("go", on_go)
    return ev


def make_honest_init(idx: int, rounds: int, out_dir: str, recv_timeout: float = 20.0):
    """Returns the init_callback of honest node ``idx`` (runs where the node runs)."""

    async def init(node):
        from examples.p2p.nodes import P2PHonestNode
        from byzpy_b200.models import SmallCNN
        from byzpy_b200.utils.data import evaluate, mnist_like, shard_indices

        worker = P2PHonestNode(indices=shard_indices(6000, N_HONEST)[idx], seed=idx)
        inbox, arrived = [], asyncio.Event()

        async def on_model(frm, payload):
            inbox.append(payload["vector"])
            arrived.set()

        node.register_message_handler("model", on_model)
        go = _go_event(node)
        app = node.application
        app.register_pipeline("half_step", _pipeline("half_step", lambda lr: worker.p2p_half_step(lr), ["lr"]))
        app.register_pipeline("aggregate", _pipeline(
            "aggregate", lambda own, received: worker.p2p_agg.aggregate([own] + list(received)),
            ["own", "received"]))
        app.register_pipeline("update_model", _pipeline(
            "update_model", lambda vector: worker.set_param_vector(vector) or True, ["vector"]))
        n_in = len(node.get_in_neighbors())

        async def training_loop():
            await go.wait()                      # every node is up (late joiners would miss broadcasts)
            for r in range(1, rounds + 1):
                own = (await node.execute_pipeline("half_step", {"lr": 0.05}))["half_step"]
                await node.broadcast_message("model", {"vector": own})
                loop = asyncio.get_running_loop()
                deadline = loop.time() + recv_timeout
                while len(inbox) < n_in and loop.time() < deadline:
                    arrived.clear()
                    try:
                        await asyncio.wait_for(arrived.wait(), timeout=max(0.01, deadline - loop.time()))
                    except asyncio.TimeoutError:
                        break
                received, inbox[:] = list(inbox), []
                agg = (await node.execute_pipeline("aggregate", {"own": own, "received": received}))["aggregate"]
                await node.execute_pipeline("update_model", {"vector": agg})
            xt, yt = mnist_like(2000, train=False)
            probe = SmallCNN()
            probe.load_state_dict(worker.dump_state_dict(), strict=True)
            loss, acc = evaluate(probe, xt, yt, torch.device("cpu"))
            with open(os.path.join(out_dir, f"honest{idx}.json"), "w") as f:
                json.dump({"node": node.node_id, "rounds": rounds, "loss": loss, "acc": acc}, f)

        await node.start_autonomous_task(training_loop(), "training_loop")

    return init


def make_byz_init(rounds: int, n_honest_in: int, recv_timeout: float = 20.0):
    async def init(node):
        from examples.p2p.nodes import P2PByzNode

        attacker = P2PByzNode()
        inbox, arrived = [], asyncio.Event()

        async def on_model(frm, payload):
            inbox.append(payload["vector"])
            arrived.set()

        node.register_message_handler("model", on_model)
        go = _go_event(node)
        node.application.register_pipeline("broadcast", _pipeline(
            "broadcast", lambda neighbor_vectors: attacker.p2p_broadcast_vector(
                neighbor_vectors=neighbor_vectors, like=neighbor_vectors[0]), ["neighbor_vectors"]))

        async def attack_loop():
            await go.wait()
            for _ in range(rounds):
                loop = asyncio.get_running_loop()
                deadline = loop.time() + recv_timeout
                while len(inbox) < n_honest_in and loop.time() < deadline:
                    arrived.clear()
                    try:
                        await asyncio.wait_for(arrived.wait(), timeout=max(0.01, deadline - loop.time()))
                    except asyncio.TimeoutError:
                        break
                seen, inbox[:] = list(inbox), []
                if seen:
                    out = (await node.execute_pipeline("broadcast", {"neighbor_vectors": seen}))["broadcast"]
                    await node.broadcast_message("model", {"vector": out})

        await node.start_autonomous_task(attack_loop(), "attack_loop")

    return init


async def main(rounds: int, context: str):
    out_dir = tempfile.mkdtemp(prefix="byzpy_b200_autonomous_")
    n = N_HONEST + N_BYZ
    topo = Topology.complete(n)
    cluster = DecentralizedCluster()
    inits = {}
    for i in range(n):
        honest = i < N_HONEST
        app_cls = HonestNodeApplication if honest else ByzantineNodeApplication
        app = app_cls(name=f"node{i}", actor_pool=[ActorPoolConfig(backend="thread", count=1)])
        ctx = ProcessContext() if context == "process" else InProcessContext()
        node = await cluster.add_node(node_id=str(i), application=app, topology=topo, context=ctx)
        init = make_honest_init(i, rounds, out_dir) if honest else make_byz_init(rounds, N_HONEST)
        if context == "process":
            node._init_callback = init          # executed inside the child once the node is up
        else:
            inits[str(i)] = init
    await cluster.start_all()
    for nid, init in inits.items():             # in-process contexts: run the callbacks here
        await init(cluster.get_node(nid))
    for i in range(n):                          # start barrier (sent by a neighbour: complete topology)
        await cluster.get_node(str((i + 1) % n)).send_message(str(i), "go", True)
    want = {f"honest{i}.json" for i in range(N_HONEST)}
    for _ in range(1200):
        if want <= set(os.listdir(out_dir)):
            break
        await asyncio.sleep(0.25)
    for name in sorted(os.listdir(out_dir)):
        print(json.load(open(os.path.join(out_dir, name))))
    await cluster.shutdown_all()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--context", default="inprocess", choices=["inprocess", "process"])
    a = ap.parse_args()
    asyncio.run(main(a.rounds, a.context))
