"""Pure-Python CPU-bound operator on an ActorPool: shows process-pool scaling where the thread backend
is GIL-bound (counterpart of the reference's benchmarks/pytorch/actor_pool_python.py).

    python benchmarks/pytorch/actor_pool_python.py --tasks 5000 --inner-iters 2000 --chunk-size 250 \
        --pool-workers 1,2,4 --pool-backend process

``--tasks`` work items of ``--inner-iters`` loop iterations each, ``--chunk-size`` items per subtask (the reference's
flags and defaults; ``--work`` is the older name of ``--inner-iters``).
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from benchmarks.pytorch._worker_args import parse_worker_counts, pool_configs  # noqa: E402
from byzpy_b200.engine.graph.operator import OpContext, Operator  # noqa: E402
from byzpy_b200.engine.graph.ops import make_single_operator_graph  # noqa: E402
from byzpy_b200.engine.graph.pool import ActorPool  # noqa: E402
from byzpy_b200.engine.graph.scheduler import NodeScheduler  # noqa: E402
from byzpy_b200.engine.graph.subtask import SubTask  # noqa: E402


def burn(seed: int, work: int) -> int:
    x = seed
    for _ in range(work):
        x = (x * 1103515245 + 12345) & 0x7FFFFFFF
    return x


def burn_many(seeds, work: int) -> int:
    return sum(burn(s, work) for s in seeds)


class BurnOp(Operator):
    name = "python-burn"
    supports_subtasks = True

    def __init__(self, work: int, chunk_size: int = 1):
        self.work, self.chunk_size = max(1, work), max(1, chunk_size)

    def compute(self, inputs, *, context: OpContext):
        return burn_many(inputs["x"], self.work)

    def create_subtasks(self, inputs, *, context: OpContext):
        seeds = list(inputs["x"])
        return [SubTask(fn=burn_many, args=(seeds[i:i + self.chunk_size], self.work), name=f"burn{i}")
                for i in range(0, len(seeds), self.chunk_size)]

    def reduce_subtasks(self, partials, inputs, *, context: OpContext):
        return sum(partials)


async def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=5000)
    ap.add_argument("--inner-iters", "--work", dest="inner_iters", type=int, default=2000)
    ap.add_argument("--chunk-size", type=int, default=250)
    ap.add_argument("--pool-workers", default="2,4,6")
    ap.add_argument("--pool-backend", default="process")
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    payload = {"x": [random.Random(a.seed + i).getrandbits(31) for i in range(a.tasks)]}
    graph = make_single_operator_graph(node_name="burn", operator=BurnOp(a.inner_iters, a.chunk_size),
                                       input_keys=("x",))
    out = {"tasks": a.tasks, "inner_iters": a.inner_iters, "chunk_size": a.chunk_size, "backend": a.pool_backend}
    t0 = time.perf_counter()
    ref = (await NodeScheduler(graph, pool=None).run(payload))["burn"]
    out["no_pool_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    for k in parse_worker_counts(a.pool_workers):
        pool = ActorPool(pool_configs(a.pool_backend, k))
        await pool.start()
        try:
            sched = NodeScheduler(graph, pool=pool)
            for _ in range(max(1, a.warmup)):
                await sched.run(payload)
            t0 = time.perf_counter()
            for _ in range(a.repeat):
                got = (await sched.run(payload))["burn"]
            out[f"pool_x{k}_ms"] = round((time.perf_counter() - t0) / a.repeat * 1e3, 2)
            assert got == ref
        finally:
            await pool.shutdown()
    print(json.dumps(out))


if __name__ == "__main__":
    asyncio.run(main())
