"""Parameter server + Multi-Krum on (synthetic) MNIST: 10 honest + 3 Byzantine SmallCNN node actors,
aggregation on the coordinator directly vs through an ActorPool (counterpart of the reference's
benchmarks/pytorch/parameter_server_actor_pool.py).

    python benchmarks/pytorch/parameter_server_actor_pool.py --rounds 50 --pool-workers 4 --pool-backend thread
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from benchmarks.pytorch._worker_args import pool_configs  # noqa: E402
from examples.ps.nodes import DistributedPSByzNode, DistributedPSHonestNode  # noqa: E402

from byzpy_b200.aggregators.geometric_wise import MultiKrum  # noqa: E402
from byzpy_b200.engine.graph.pool import ActorPool  # noqa: E402
from byzpy_b200.engine.node.actors import ByzantineNodeActor, HonestNodeActor  # noqa: E402
from byzpy_b200.engine.parameter_server.ps import ParameterServer  # noqa: E402
from byzpy_b200.utils.data import shard_indices  # noqa: E402


async def run(rounds, n_h, n_b, pool):
    shards = shard_indices(6000, n_h)
    hon = [await HonestNodeActor.spawn(DistributedPSHonestNode, backend="thread",
                                       kwargs=dict(indices=shards[i], seed=i)) for i in range(n_h)]
    byz = [await ByzantineNodeActor.spawn(DistributedPSByzNode, backend="thread") for _ in range(n_b)]
    ps = ParameterServer(hon, byz, MultiKrum(f=n_b, q=n_h - n_b), actor_pool=pool)
    await ps.round()                         # warm-up
    t0 = time.perf_counter()
    for _ in range(rounds):
        await ps.round()
    dt = (time.perf_counter() - t0) / rounds * 1e3
    await ps.shutdown()
    return dt


async def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=50)
    ap.add_argument("--honest", type=int, default=10)
    ap.add_argument("--byzantine", type=int, default=3)
    ap.add_argument("--pool-workers", type=int, default=4)
    ap.add_argument("--pool-backend", default="thread")
    a = ap.parse_args()
    out = {"rounds": a.rounds, "honest": a.honest, "byzantine": a.byzantine}
    out["direct_ms_per_round"] = round(await run(a.rounds, a.honest, a.byzantine, None), 2)
    pool = ActorPool(pool_configs(a.pool_backend, a.pool_workers))
    await pool.start()
    try:
        out[f"pool_x{a.pool_workers}_ms_per_round"] = round(await run(a.rounds, a.honest, a.byzantine, pool), 2)
    finally:
        await pool.shutdown()
    print(json.dumps(out))


if __name__ == "__main__":
    asyncio.run(main())
