"""NodeScheduler vs ParallelScheduler on k independent branches ``preprocess -> median`` (the
reference's benchmarks/scheduler/pipeline_benchmark.py).  On CUDA inputs the parallel scheduler
issues each branch on its own CUDA stream.

    python benchmarks/scheduler/pipeline_benchmark.py --branches 4 --num-grads 64 --grad-dim 200000
"""
from __future__ import annotations

import argparse
import asyncio
import functools
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian  # noqa: E402
from byzpy_b200.engine.graph.graph import ComputationGraph, GraphNode, graph_input  # noqa: E402
from byzpy_b200.engine.graph.ops import CallableOp  # noqa: E402
from byzpy_b200.engine.graph.parallel_scheduler import ParallelScheduler  # noqa: E402
from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig  # noqa: E402
from byzpy_b200.engine.graph.scheduler import NodeScheduler  # noqa: E402


def preprocess(vectors, iters=30):
    out = list(vectors)
    for _ in range(iters):
        out = [v * 0.999 + 0.001 for v in out]
    return out


async def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--branches", type=int, default=4)
    ap.add_argument("--num-grads", type=int, default=64)
    ap.add_argument("--grad-dim", type=int, default=200000)
    ap.add_argument("--chunk-size", type=int, default=8192, help="the median's subtask granularity")
    ap.add_argument("--pool-workers", default="0",
                    help="worker counts, comma or space separated (reference default 2,4,6); 0 = no pool")
    ap.add_argument("--pool-backend", default="thread")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--preprocess-iterations", type=int, default=30)
    ap.add_argument("--max-pending-subtasks", type=int, default=None,
                    help="cap on subtasks in flight over all concurrent operators (default: 8 x pool size)")
    a = ap.parse_args()
    dev = torch.device(a.device)
    gen = torch.Generator().manual_seed(a.seed)
    data = [torch.randn(a.grad_dim, generator=gen).to(dev) for _ in range(a.num_grads)]
    pre = functools.partial(preprocess, iters=a.preprocess_iterations)
    nodes = []
    for b in range(a.branches):
        nodes.append(GraphNode(f"pre{b}", CallableOp(pre, input_mapping={"vectors": "vectors"}),
                               {"vectors": graph_input("vectors")}))
        nodes.append(GraphNode(f"med{b}", CoordinateWiseMedian(chunk_size=a.chunk_size), {"gradients": f"pre{b}"}))
    graph = ComputationGraph(nodes, outputs=[f"med{b}" for b in range(a.branches)])
    sync = (lambda: torch.cuda.synchronize()) if dev.type == "cuda" else (lambda: None)
    counts = [int(k) for k in str(a.pool_workers).replace(",", " ").split()] or [0]
    for workers in counts:
        pool = None
        if workers:
            pool = ActorPool([ActorPoolConfig(backend=a.pool_backend, count=workers)])
            await pool.start()
        res = {"branches": a.branches, "n": a.num_grads, "d": a.grad_dim, "device": a.device, "pool": workers}
        try:
            for name, sched in (("node_scheduler_ms", NodeScheduler(graph, pool=pool)),
                                ("parallel_scheduler_ms",
                                 ParallelScheduler(graph, pool=pool, max_pending_subtasks=a.max_pending_subtasks))):
                for _ in range(max(1, a.warmup)):
                    await sched.run({"vectors": data})
                sync()
                t0 = time.perf_counter()
                for _ in range(a.repeat):
                    await sched.run({"vectors": data})
                sync()
                res[name] = round((time.perf_counter() - t0) / a.repeat * 1e3, 2)
        finally:
            if pool is not None:
                await pool.shutdown()
        res["speedup"] = round(res["node_scheduler_ms"] / res["parallel_scheduler_ms"], 2)
        print(json.dumps(res))


if __name__ == "__main__":
    asyncio.run(main())
