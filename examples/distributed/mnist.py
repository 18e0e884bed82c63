"""Multi-machine parameter server over TCP actor servers (counterpart of the reference's
examples/distributed/mnist.py, same flags): node actors are placed round-robin on the listed servers,
Multi-Krum aggregation on the coordinator, periodic evaluation of honest node 0's model.

    # on every worker machine:   python examples/distributed/server.py --port 29000
    python examples/distributed/mnist.py --remote-hosts tcp://10.0.0.2:29000,tcp://10.0.0.3:29000 --rounds 100
    # single-box smoke test (starts two loopback servers itself):
    python examples/distributed/mnist.py --local --rounds 5

``--servers host:port,...`` is the older spelling of ``--remote-hosts``.
"""
from __future__ import annotations

import argparse
import asyncio
import os
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from examples.ps.nodes import DistributedPSByzNode, DistributedPSHonestNode, SmallCNN  # noqa: E402

from byzpy_b200.aggregators.geometric_wise import MultiKrum  # noqa: E402
from byzpy_b200.engine.actor.backends.remote import RemoteActorServer  # noqa: E402
from byzpy_b200.engine.node.actors import ByzantineNodeActor, HonestNodeActor  # noqa: E402
from byzpy_b200.engine.parameter_server.ps import ParameterServer  # noqa: E402
from byzpy_b200.utils.data import evaluate, mnist_like, shard_indices  # noqa: E402


def _spec(host: str) -> str:
    return host if "://" in host else f"tcp://{host}"


async def main(a):
    owned = []
    hosts = [_spec(h.strip()) for h in (a.remote_hosts or a.servers or "").split(",") if h.strip()]
    if a.local or not hosts:
        for _ in range(2):
            srv = RemoteActorServer("127.0.0.1", 0)
            await srv.start()
            asyncio.ensure_future(srv._server.serve_forever())
            owned.append(srv)
        hosts = [f"tcp://127.0.0.1:{s.port}" for s in owned]
    n = a.num_honest + a.num_byz
    if len(hosts) < n:
        print(f"{len(hosts)} hosts for {n} nodes: placing nodes round-robin", file=sys.stderr)
    q = max(1, n - a.f - 1) if a.q is None else a.q
    torch.manual_seed(a.seed)
    print(f"honest {a.num_honest}  byzantine {a.num_byz}  hosts {hosts}  rounds {a.rounds}  Multi-Krum f={a.f} q={q}")
    shards = shard_indices(6000, a.num_honest)
    hon = [await HonestNodeActor.spawn(DistributedPSHonestNode, backend=hosts[i % len(hosts)],
                                       kwargs=dict(indices=shards[i], seed=a.seed + i, batch_size=a.batch_size, lr=a.lr,
                                                   data_root=a.data_root)) for i in range(a.num_honest)]
    byz = [await ByzantineNodeActor.spawn(DistributedPSByzNode, backend=hosts[(a.num_honest + j) % len(hosts)])
           for j in range(a.num_byz)]
    ps = ParameterServer(hon, byz, MultiKrum(f=a.f, q=q, chunk_size=a.chunk_size), node_timeout=60.0,
                         tolerate_failures=True)
    eval_model = SmallCNN()
    xt, yt = mnist_like(2000, train=False, root=a.data_root)

    async def report(tag, node):
        eval_model.load_state_dict(await node.dump_state_dict())
        loss, acc = await asyncio.to_thread(evaluate, eval_model, xt, yt, torch.device("cpu"))
        print(f"{tag} test loss={loss:.4f}  acc={acc:.4f}  elapsed={time.perf_counter() - t0:.2f}s  "
              f"failed so far: {len(ps.failed)}")

    t0 = time.perf_counter()
    for r in range(1, a.rounds + 1):
        await ps.round()
        if a.eval_interval > 0 and r % a.eval_interval == 0:
            await report(f"[round {r:04d}]", hon[0])
    total = time.perf_counter() - t0
    lost = {rec[1] for rec in ps.failed}            # "honest:<i>" / "byz:<j>" of nodes that stopped answering
    for i, h in enumerate(hon):
        if f"honest:{i}" not in lost:
            await report(f"node {i} ({hosts[i % len(hosts)]}):", h)
    print(f"total {total:.2f}s, {total / max(1, a.rounds):.3f}s per round")
    await ps.shutdown()
    for s in owned:
        await s.stop()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--remote-hosts", default="", help="tcp://host:port,... of running actor servers")
    ap.add_argument("--servers", default="", help="older spelling of --remote-hosts (host:port,...)")
    ap.add_argument("--local", action="store_true", help="start two loopback actor servers in this process")
    ap.add_argument("--num-honest", type=int, default=3)
    ap.add_argument("--num-byz", type=int, default=1)
    ap.add_argument("--rounds", type=int, default=50)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--lr", type=float, default=0.05)
    ap.add_argument("--f", type=int, default=1)
    ap.add_argument("--q", type=int, default=None, help="Multi-Krum q (default n - f - 1)")
    ap.add_argument("--chunk-size", type=int, default=32)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--data-root", default="./data")
    ap.add_argument("--eval-interval", type=int, default=10, help="evaluate every N rounds (0: never)")
    asyncio.run(main(ap.parse_args()))
