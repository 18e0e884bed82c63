"""Parameter-server training with node actors hosted by remote TCP actor servers (``tcp://host:port`` backends):
4 honest SmallCNN workers + 1 Empire Byzantine worker, CoordinateWiseMedian on the coordinator.

    # on each worker host:  python examples/ps/remote/server.py --port 29000
    python examples/ps/remote/mnist.py --servers 10.0.0.2:29000,10.0.0.3:29000
    python examples/ps/remote/mnist.py --local           # single-box smoke test
"""
from __future__ import annotations

import argparse
import asyncio
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from examples.ps.nodes import DistributedPSByzNode, DistributedPSHonestNode  # noqa: E402

from byzpy_b200.engine.actor.backends.remote import RemoteActorServer  # noqa: E402
from byzpy_b200.engine.node.actors import ByzantineNodeActor, HonestNodeActor  # noqa: E402
from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian  # noqa: E402
from byzpy_b200.engine.parameter_server.ps import ParameterServer  # noqa: E402
from byzpy_b200.models import SmallCNN  # noqa: E402
from byzpy_b200.utils.data import evaluate, mnist_like, shard_indices  # noqa: E402


async def main(rounds: int, servers, local: bool):
    owned = []
    if local:       # single-box smoke test: host the actor server in this process
        srv = RemoteActorServer("127.0.0.1", 0)
        await srv.start()
        asyncio.ensure_future(srv._server.serve_forever())
        owned.append(srv)
        servers = [f"127.0.0.1:{srv.port}"]
    backends = [f"tcp://{s}" for s in servers]
    n_h, n_b = 4, 1
    shards = shard_indices(6000, n_h)
    hon = [await HonestNodeActor.spawn(DistributedPSHonestNode, backend=backends[i % len(backends)],
                                       kwargs=dict(indices=shards[i], seed=i)) for i in range(n_h)]
    byz = [await ByzantineNodeActor.spawn(DistributedPSByzNode, backend=backends[(n_h + i) % len(backends)])
           for i in range(n_b)]
    ps = ParameterServer(hon, byz, CoordinateWiseMedian(), node_timeout=60.0, tolerate_failures=True)
    xt, yt = mnist_like(2000, train=False)
    probe = SmallCNN()
    for r in range(1, rounds + 1):
        await ps.round()
        if r % max(1, rounds // 5) == 0:
            probe.load_state_dict(await hon[0].dump_state_dict(), strict=True)
            loss, acc = evaluate(probe, xt, yt, torch.device("cpu"))
            print(f"[round {r:04d}] node0 test loss={loss:.4f} acc={acc:.4f}")
    await ps.shutdown()
    for s in owned:
        await s.stop()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=50)
    ap.add_argument("--servers", default="", help="comma separated host:port of actor servers")
    ap.add_argument("--local", action="store_true", help="start one actor server in-process")
    a = ap.parse_args()
    asyncio.run(main(a.rounds, [s for s in a.servers.split(",") if s], a.local or not a.servers))
