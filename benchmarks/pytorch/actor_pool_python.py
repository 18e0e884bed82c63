"""Pure-Python CPU-bound operator on an ActorPool: shows process-pool scaling where the thread backend
is GIL-bound (counterpart of the reference's benchmarks/pytorch/actor_pool_python.py).

    python benchmarks/pytorch/actor_pool_python.py --tasks 16 --work 200000 --pool-workers 1,2,4 --pool-backend process
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
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


class BurnOp(Operator):
    name = "python-burn"
    supports_subtasks = True

    def __init__(self, tasks: int, work: int):
        self.tasks, self.work = tasks, work

    def compute(self, inputs, *, context: OpContext):
        return sum(burn(i, self.work) for i in range(self.tasks))

    def create_subtasks(self, inputs, *, context: OpContext):
        return [SubTask(fn=burn, args=(i, self.work), name=f"burn{i}") for i in range(self.tasks)]

    def reduce_subtasks(self, partials, inputs, *, context: OpContext):
        return sum(partials)


async def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=16)
    ap.add_argument("--work", type=int, default=200000)
    ap.add_argument("--pool-workers", default="1,2,4")
    ap.add_argument("--pool-backend", default="process")
    ap.add_argument("--repeat", type=int, default=2)
    a = ap.parse_args()
    graph = make_single_operator_graph(node_name="burn", operator=BurnOp(a.tasks, a.work), input_keys=("x",))
    out = {"tasks": a.tasks, "work": a.work, "backend": a.pool_backend}
    t0 = time.perf_counter()
    ref = (await NodeScheduler(graph, pool=None).run({"x": 0}))["burn"]
    out["no_pool_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    for k in parse_worker_counts(a.pool_workers):
        pool = ActorPool(pool_configs(a.pool_backend, k))
        await pool.start()
        try:
            sched = NodeScheduler(graph, pool=pool)
            await sched.run({"x": 0})
            t0 = time.perf_counter()
            for _ in range(a.repeat):
                got = (await sched.run({"x": 0}))["burn"]
            out[f"pool_x{k}_ms"] = round((time.perf_counter() - t0) / a.repeat * 1e3, 2)
            assert got == ref
        finally:
            await pool.shutdown()
    print(json.dumps(out))


if __name__ == "__main__":
    asyncio.run(main())
