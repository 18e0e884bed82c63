"""Tour of the actor layer: any class as an actor on any backend, cross-backend channels, an
ActorPool running operator subtasks, a lazy graph on the ParallelScheduler.

    python examples/actor_demo/actor_demo.py
"""
from __future__ import annotations

import asyncio
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

from byzpy_b200 import run_operator  # noqa: E402
from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian  # noqa: E402
from byzpy_b200.aggregators.geometric_wise import MultiKrum  # noqa: E402
from byzpy_b200.engine.actor.base import ActorRef  # noqa: E402
from byzpy_b200.engine.actor.factory import resolve_backend  # noqa: E402
from byzpy_b200.engine.graph.lazy import GraphBuilder  # noqa: E402
from byzpy_b200.engine.graph.parallel_scheduler import ParallelScheduler  # noqa: E402
from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig  # noqa: E402
from byzpy_b200.pre_aggregators import Clipping  # noqa: E402


class Accumulator:
    def __init__(self, start=0.0):
        self.total = start

    def add(self, x):
        self.total += float(x)
        return self.total


async def main():
    # 1) the same class on three backends
    for spec in ("thread", "gpu", "process"):
        be = resolve_backend(spec)
        async with ActorRef(be) as ref:
            await be.construct(Accumulator, args=(), kwargs={"start": 1.0})
            print(f"{spec:8s} actor -> {await ref.add(41.0)}")
    # 2) a channel from a thread actor to a gpu (CUDA-stream) actor
    a, b = resolve_backend("thread"), resolve_backend("gpu")
    for be in (a, b):
        await be.start()
        await be.construct(Accumulator, args=(), kwargs={})
    ra, rb = ActorRef(a), ActorRef(b)
    ca, cb = await ra.open_channel("demo"), await rb.open_channel("demo")
    await ca.send(await rb.endpoint(), {"grad": torch.ones(4)})
    print("gpu actor received:", (await cb.recv(timeout=1.0))["grad"])
    await a.close()
    await b.close()
    # 3) operator subtasks on a heterogeneous pool
    grads = [torch.randn(100_000) for _ in range(16)]
    out = await run_operator(CoordinateWiseMedian(), {"gradients": grads},
                             pool_config=[ActorPoolConfig("thread", count=2), ActorPoolConfig("gpu", count=2)])
    print("pooled median == direct:", torch.equal(out, CoordinateWiseMedian().aggregate(grads)))
    # 4) a lazy two-branch graph on the dataflow scheduler
    b_ = GraphBuilder()
    x = b_.input("vectors")
    clipped = x.apply(Clipping(threshold=300.0))
    med = clipped.apply(CoordinateWiseMedian(), name="median")
    krum = clipped.apply(MultiKrum(f=3, q=5), name="krum")
    res = await ParallelScheduler(b_.build(outputs=[med.key, krum.key])).run({"vectors": grads})
    print("graph outputs:", {k: tuple(v.shape) for k, v in res.items()})


if __name__ == "__main__":
    asyncio.run(main())
