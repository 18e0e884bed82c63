"""BASELINE.json config 1: CoordinateWiseMedian.aggregate on 10 x torch.randn(1000), CPU, direct and
through NodeScheduler + ActorPool(thread x 4) -- plumbing latency, no GPU.  With --reference the same
calls go through the unmodified reference package from baseline/_ref.

    python benchmarks/config1_cpu_plumbing.py [--reference]
"""
import argparse
import asyncio
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--reference", action="store_true")
ap.add_argument("--repeat", type=int, default=50)
a = ap.parse_args()
if a.reference:
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    from byzpy.aggregators.coordinate_wise import CoordinateWiseMedian
    from byzpy.engine.graph.ops import make_single_operator_graph
    from byzpy.engine.graph.pool import ActorPool, ActorPoolConfig
    from byzpy.engine.graph.scheduler import NodeScheduler
else:
    sys.path.insert(0, ROOT)
    from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
    from byzpy_b200.engine.graph.ops import make_single_operator_graph
    from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig
    from byzpy_b200.engine.graph.scheduler import NodeScheduler

torch.manual_seed(0)
grads = [torch.randn(1000) for _ in range(10)]


def timed(fn, n):
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


async def main():
    out = {"impl": "reference" if a.reference else "ours"}
    t0 = time.perf_counter()
    CoordinateWiseMedian().aggregate(grads)
    out["direct_first_call_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
    out["direct_ms"] = round(timed(lambda: CoordinateWiseMedian().aggregate(grads), a.repeat), 4)
    graph = make_single_operator_graph(node_name="agg", operator=CoordinateWiseMedian(), input_keys=("gradients",))
    pool = ActorPool([ActorPoolConfig(backend="thread", count=4)])
    await pool.start()
    sched = NodeScheduler(graph, pool=pool)
    t0 = time.perf_counter()
    await sched.run({"gradients": grads})
    out["pool_first_call_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
    t0 = time.perf_counter()
    for _ in range(a.repeat):
        await sched.run({"gradients": grads})
    out["pool_x4_ms"] = round((time.perf_counter() - t0) / a.repeat * 1e3, 4)
    await pool.shutdown()
    print(json.dumps(out))


asyncio.run(main())
