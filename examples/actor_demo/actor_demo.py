"""Tour of the actor layer: any class as an actor on any backend, cross-backend channels, an
ActorPool running operator subtasks, a lazy graph on the ParallelScheduler.

    python examples/actor_demo/actor_demo.py                       # every local demo
    python examples/actor_demo/actor_demo.py --demo channel        # one of: all basic channel pipeline
                                                                   #         channel-detail pool graph remote
    python examples/actor_demo/remote_server.py --port 29000 &     # then:
    python examples/actor_demo/actor_demo.py --with-remote --remote-host 127.0.0.1 --remote-port 29000

(``--demo`` / ``--with-remote`` / ``--remote-host`` / ``--remote-port`` as in the reference's
examples/actor_demo/actor_demo.py:574-597.)
"""
from __future__ import annotations

import argparse
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


class Stage:
    """One stage of the pipeline demo: squares what it is given and counts its calls."""

    def __init__(self, name):
        self.name, self.calls = name, 0

    def work(self, x):
        self.calls += 1
        return float(x) ** 2

    def info(self):
        return f"{self.name}: {self.calls} calls"


async def demo_basic():
    """The same class on three backends; attribute access on the reference is an awaitable remote call."""
    for spec in ("thread", "gpu", "process"):
        be = resolve_backend(spec)
        async with ActorRef(be) as ref:
            await be.construct(Accumulator, args=(), kwargs={"start": 1.0})
            print(f"{spec:8s} actor -> {await ref.add(41.0)}")


async def demo_channel():
    """A channel from a thread actor to a gpu (CUDA-stream) actor."""
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


async def demo_pipeline():
    """Three thread actors compute concurrently, one process actor collects: data crosses backends by value."""
    stages = []
    for i in range(3):
        be = resolve_backend("thread")
        await be.start()
        await be.construct(Stage, args=(f"stage-{i}",), kwargs={})
        stages.append(ActorRef(be))
    sink_be = resolve_backend("process")
    await sink_be.start()
    await sink_be.construct(Accumulator, args=(), kwargs={})
    sink = ActorRef(sink_be)
    squares = await asyncio.gather(*(st.work(i + 1) for i, st in enumerate(stages)))
    for v in squares:
        total = await sink.add(v)
    print("pipeline: squares", squares, "-> running total on the process actor:", total)
    print("pipeline:", [await st.info() for st in stages])
    for st in stages:
        await st._backend.close()
    await sink_be.close()


async def demo_channel_detail():
    """Producer / consumer over a named mailbox, thread -> process and back: ordering, timeouts, endpoints."""
    prod_be, cons_be = resolve_backend("thread"), resolve_backend("process")
    for be in (prod_be, cons_be):
        await be.start()
        await be.construct(Accumulator, args=(), kwargs={})
    prod, cons = ActorRef(prod_be), ActorRef(cons_be)
    out, inbox = await prod.open_channel("work"), await cons.open_channel("work")
    to_consumer, to_producer = await cons.endpoint(), await prod.endpoint()
    print("endpoints:", to_producer, "->", to_consumer)
    for i in range(3):
        await out.send(to_consumer, {"seq": i, "payload": torch.full((2,), float(i))})
    got = [await inbox.recv(timeout=2.0) for _ in range(3)]
    print("consumer received in order:", [m["seq"] for m in got])
    await inbox.send(to_producer, {"ack": len(got)})
    print("producer received:", await out.recv(timeout=2.0))
    print("empty mailbox: recv(timeout=0.05) ->", await inbox.recv(timeout=0.05))
    await prod_be.close()
    await cons_be.close()


async def demo_pool(grads):
    """Operator subtasks on a heterogeneous pool."""
    out = await run_operator(CoordinateWiseMedian(), {"gradients": grads},
                             pool_config=[ActorPoolConfig("thread", count=2), ActorPoolConfig("gpu", count=2)])
    print("pooled median == direct:", torch.equal(out, CoordinateWiseMedian().aggregate(grads)))


async def demo_graph(grads):
    """A lazy two-branch graph on the dataflow scheduler."""
    b_ = GraphBuilder()
    x = b_.input("vectors")
    clipped = x.apply(Clipping(threshold=300.0))
    med = clipped.apply(CoordinateWiseMedian(), name="median")
    krum = clipped.apply(MultiKrum(f=3, q=5), name="krum")
    res = await ParallelScheduler(b_.build(outputs=[med.key, krum.key])).run({"vectors": grads})
    print("graph outputs:", {k: tuple(v.shape) for k, v in res.items()})


async def demo_remote(host, port):
    """The same calls against an actor living in ``remote_server.py``: only the backend spec differs."""
    be = resolve_backend(f"tcp://{host}:{port}")
    try:
        await be.start()
    except OSError as exc:
        print(f"remote demo: no actor server at {host}:{port} ({exc}); start one with\n"
              f"    python examples/actor_demo/remote_server.py --port {port}")
        return
    await be.construct(Accumulator, args=(), kwargs={"start": 1.0})
    remote = ActorRef(be)
    local_be = resolve_backend("thread")
    async with ActorRef(local_be) as local:
        await local_be.construct(Accumulator, args=(), kwargs={"start": 1.0})
        print("local  (thread):", await local.add(41.0))
        print(f"remote (tcp://{host}:{port}):", await remote.add(41.0))
    await be.close()


async def main(a):
    grads = [torch.randn(100_000) for _ in range(16)]
    demos = {"basic": demo_basic, "channel": demo_channel, "pipeline": demo_pipeline,
             "channel-detail": demo_channel_detail, "pool": lambda: demo_pool(grads), "graph": lambda: demo_graph(grads)}
    if a.demo == "remote":
        return await demo_remote(a.remote_host, a.remote_port)
    for name, fn in demos.items():
        if a.demo in ("all", name):
            print(f"--- {name}")
            await fn()
    if a.with_remote and a.demo == "all":
        print("--- remote")
        await demo_remote(a.remote_host, a.remote_port)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--demo", default="all",
                    choices=["all", "basic", "channel", "pipeline", "channel-detail", "pool", "graph", "remote"])
    ap.add_argument("--with-remote", action="store_true", help="also run the remote demo (needs remote_server.py)")
    ap.add_argument("--remote-host", default="localhost")
    ap.add_argument("--remote-port", type=int, default=29000)
    asyncio.run(main(ap.parse_args()))
