"""Multi-machine parameter server over TCP actor servers (counterpart of the reference's
examples/distributed/mnist.py): node actors are placed round-robin on the listed servers,
Multi-Krum aggregation on the coordinator.

    # on every worker machine:   python examples/distributed/server.py --port 29000
    python examples/distributed/mnist.py --servers 10.0.0.2:29000,10.0.0.3:29000 --rounds 100
    # single-box smoke test:     python examples/distributed/mnist.py --local
"""
from __future__ import annotations

import argparse
import asyncio
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from examples.ps.nodes import DistributedPSByzNode, DistributedPSHonestNode  # noqa: E402

from byzpy_b200.aggregators.geometric_wise import MultiKrum  # noqa: E402
from byzpy_b200.engine.actor.backends.remote import RemoteActorServer  # noqa: E402
from byzpy_b200.engine.node.actors import ByzantineNodeActor, HonestNodeActor  # noqa: E402
from byzpy_b200.engine.parameter_server.ps import ParameterServer  # noqa: E402
from byzpy_b200.utils.data import shard_indices  # noqa: E402


async def main(servers, rounds, local):
    owned = []
    if local:
        for _ in range(2):
            srv = RemoteActorServer("127.0.0.1", 0)
            await srv.start()
            asyncio.ensure_future(srv._server.serve_forever())
            owned.append(srv)
        servers = [f"127.0.0.1:{s.port}" for s in owned]
    n_h, n_b = 6, 2
    shards = shard_indices(6000, n_h)
    hon = [await HonestNodeActor.spawn(DistributedPSHonestNode, backend=f"tcp://{servers[i % len(servers)]}",
                                       kwargs=dict(indices=shards[i], seed=i)) for i in range(n_h)]
    byz = [await ByzantineNodeActor.spawn(DistributedPSByzNode, backend=f"tcp://{servers[i % len(servers)]}")
           for i in range(n_b)]
    ps = ParameterServer(hon, byz, MultiKrum(f=n_b, q=n_h - 1), node_timeout=60.0, tolerate_failures=True)
    for r in range(1, rounds + 1):
        g = await ps.round()
        if r % max(1, rounds // 5) == 0:
            print(f"[round {r:04d}] |aggregate| = {g.norm().item():.4f}  failed so far: {len(ps.failed)}")
    await ps.shutdown()
    for s in owned:
        await s.stop()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--servers", default="")
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--local", action="store_true")
    a = ap.parse_args()
    asyncio.run(main([s for s in a.servers.split(",") if s], a.rounds, a.local or not a.servers))
