"""Parameter server + Multi-Krum on (synthetic) MNIST: honest + Byzantine SmallCNN node actors, aggregation on the
coordinator directly vs through an ActorPool of each size in ``--pool-workers`` (counterpart of the reference's
benchmarks/pytorch/parameter_server_actor_pool.py, same flags and defaults; ``--honest`` / ``--byzantine`` are kept as
aliases of ``--num-honest`` / ``--num-byz``).

    python benchmarks/pytorch/parameter_server_actor_pool.py --rounds 50 --pool-workers 2,4 --pool-backend thread \
        --actor-backend thread
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch  # noqa: E402

from benchmarks.pytorch._worker_args import DEFAULT_WORKER_COUNTS, coerce_worker_counts, pool_configs  # noqa: E402
from examples.ps.nodes import DistributedPSByzNode, DistributedPSHonestNode  # noqa: E402

from byzpy_b200.aggregators.geometric_wise import MultiKrum  # noqa: E402
from byzpy_b200.engine.graph.pool import ActorPool  # noqa: E402
from byzpy_b200.engine.node.actors import ByzantineNodeActor, HonestNodeActor  # noqa: E402
from byzpy_b200.engine.parameter_server.ps import ParameterServer  # noqa: E402
from byzpy_b200.utils.data import shard_indices  # noqa: E402


async def run(a, pool):
    torch.manual_seed(a.seed)
    shards = shard_indices(6000, a.num_honest)
    hon = [await HonestNodeActor.spawn(DistributedPSHonestNode, backend=a.actor_backend,
                                       kwargs=dict(indices=shards[i], seed=a.seed + i, batch_size=a.batch_size,
                                                   lr=a.lr, data_root=a.data_root))
           for i in range(a.num_honest)]
    byz = [await ByzantineNodeActor.spawn(DistributedPSByzNode, backend=a.actor_backend) for _ in range(a.num_byz)]
    ps = ParameterServer(hon, byz, MultiKrum(f=a.f, q=a.q, chunk_size=a.chunk_size), actor_pool=pool)
    await ps.round()                         # warm-up
    t0 = time.perf_counter()
    for _ in range(a.rounds):
        await ps.round()
    dt = (time.perf_counter() - t0) / a.rounds * 1e3
    await ps.shutdown()
    return dt


async def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-honest", "--honest", dest="num_honest", type=int, default=10)
    ap.add_argument("--num-byz", "--byzantine", dest="num_byz", type=int, default=3)
    ap.add_argument("--rounds", type=int, default=50)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--lr", type=float, default=0.1)
    ap.add_argument("--f", type=int, default=None, help="Multi-Krum f (default: --num-byz)")
    ap.add_argument("--q", type=int, default=None, help="Multi-Krum q (default: n - f - 1)")
    ap.add_argument("--chunk-size", type=int, default=32)
    ap.add_argument("--pool-workers", default=",".join(map(str, DEFAULT_WORKER_COUNTS)))
    ap.add_argument("--pool-backend", default="process")
    ap.add_argument("--actor-backend", default="process")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--data-root", default="./data")
    a = ap.parse_args()
    n = a.num_honest + a.num_byz
    a.f = a.num_byz if a.f is None else a.f
    a.q = max(1, n - a.f - 1) if a.q is None else a.q
    out = {"rounds": a.rounds, "honest": a.num_honest, "byzantine": a.num_byz, "f": a.f, "q": a.q}
    out["direct_ms_per_round"] = round(await run(a, None), 2)
    for k in coerce_worker_counts(a.pool_workers):
        pool = ActorPool(pool_configs(a.pool_backend, k))
        await pool.start()
        try:
            out[f"pool_x{k}_ms_per_round"] = round(await run(a, pool), 2)
        finally:
            await pool.shutdown()
    print(json.dumps(out))


if __name__ == "__main__":
    asyncio.run(main())
