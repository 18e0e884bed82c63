"""Operator benchmark: direct call vs NodeScheduler(pool=None) vs ActorPool x k workers -- the
shape of every reference script under benchmarks/pytorch/*_actor_pool.py, for any operator.

    python benchmarks/operator_pool_bench.py --op median --num-grads 64 --grad-dim 65536 \
        --pool-workers 2,4,6 --pool-backend thread --repeat 3
    python benchmarks/operator_pool_bench.py --op median --num-grads 10 --grad-dim 1000   # BASELINE config 1
    python benchmarks/operator_pool_bench.py --op multi-krum --device cuda --pool-backend gpu
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean, MeanOfMedians  # noqa: E402
from byzpy_b200.aggregators.geometric_wise import (SMEA, GeometricMedian, Krum, MinimumDiameterAveraging,  # noqa: E402
                                                    MoNNA, MultiKrum)
from byzpy_b200.aggregators.norm_wise import CAF, CenteredClipping, ComparativeGradientElimination  # noqa: E402
from byzpy_b200.attacks import (EmpireAttack, GaussianAttack, InfAttack, LittleAttack, MimicAttack,  # noqa: E402
                                SignFlipAttack)
from byzpy_b200.engine.graph.ops import make_single_operator_graph  # noqa: E402
from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig  # noqa: E402
from byzpy_b200.engine.graph.scheduler import NodeScheduler  # noqa: E402
from byzpy_b200.pre_aggregators import ARC, Bucketing, Clipping, NearestNeighborMixing  # noqa: E402


def make(op: str, n: int, f: int):
    table = {
        "median": (lambda: CoordinateWiseMedian(), "gradients"),
        "trimmed-mean": (lambda: CoordinateWiseTrimmedMean(f=f), "gradients"),
        "meamed": (lambda: MeanOfMedians(f=f), "gradients"),
        "multi-krum": (lambda: MultiKrum(f=f, q=max(1, n - 2 * f)), "gradients"),
        "krum": (lambda: Krum(f=f), "gradients"),
        "geometric-median": (lambda: GeometricMedian(), "gradients"),
        "mda": (lambda: MinimumDiameterAveraging(f=f), "gradients"),
        "monna": (lambda: MoNNA(f=f), "gradients"),
        "smea": (lambda: SMEA(f=f), "gradients"),
        "centered-clipping": (lambda: CenteredClipping(c_tau=0.1, M=10), "gradients"),
        "cge": (lambda: ComparativeGradientElimination(f=f), "gradients"),
        "caf": (lambda: CAF(f=f), "gradients"),
        "clipping": (lambda: Clipping(threshold=2.0), "vectors"),
        "arc": (lambda: ARC(f=f), "vectors"),
        "nnm": (lambda: NearestNeighborMixing(f=f), "vectors"),
        "bucketing": (lambda: Bucketing(bucket_size=max(1, n // 16)), "vectors"),
        "empire": (lambda: EmpireAttack(), "honest_grads"),
        "little": (lambda: LittleAttack(f=f), "honest_grads"),
        "gaussian": (lambda: GaussianAttack(seed=0), "honest_grads"),
        "inf": (lambda: InfAttack(), "honest_grads"),
        "mimic": (lambda: MimicAttack(epsilon=0), "honest_grads"),
        "sign-flip": (lambda: SignFlipAttack(scale=-1.0), "base_grad"),
    }
    if op not in table:
        raise SystemExit(f"unknown --op {op!r}; choose from {sorted(table)}")
    return table[op]


def direct_call(operator, key, data):
    if key == "gradients":
        return operator.aggregate(data)
    if key == "vectors":
        return operator.pre_aggregate(data)
    if key == "base_grad":
        return operator.apply(base_grad=data)
    return operator.apply(honest_grads=data)


def timed(fn, warmup, repeat, sync):
    for _ in range(warmup):
        fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(repeat):
        fn()
    sync()
    return (time.perf_counter() - t0) / repeat * 1e3


async def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", default="median")
    ap.add_argument("--num-grads", type=int, default=64)
    ap.add_argument("--grad-dim", type=int, default=65536)
    ap.add_argument("--f", type=int, default=8)
    ap.add_argument("--pool-workers", default="2,4,6")
    ap.add_argument("--pool-backend", default="thread")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device(a.device)
    g = torch.Generator().manual_seed(a.seed)
    data = [torch.randn(a.grad_dim, generator=g).to(dev) for _ in range(a.num_grads)]
    f = min(a.f, max(0, (a.num_grads - 1) // 2 - 1))
    mk, key = make(a.op, a.num_grads, f)
    if key == "base_grad":  # SignFlip consumes one vector (the Byzantine node's own gradient)
        data = data[0]
    sync = (lambda: torch.cuda.synchronize()) if dev.type == "cuda" else (lambda: None)
    out = {"op": a.op, "n": a.num_grads, "d": a.grad_dim, "device": a.device, "backend": a.pool_backend}
    out["direct_ms"] = round(timed(lambda: direct_call(mk(), key, data), a.warmup, a.repeat, sync), 3)
    graph = make_single_operator_graph(node_name="op", operator=mk(), input_keys=(key,))

    async def via(pool):
        sched = NodeScheduler(graph, pool=pool)
        for _ in range(a.warmup):
            await sched.run({key: data})
        sync()
        t0 = time.perf_counter()
        for _ in range(a.repeat):
            await sched.run({key: data})
        sync()
        return (time.perf_counter() - t0) / a.repeat * 1e3

    out["scheduler_no_pool_ms"] = round(await via(None), 3)
    for k in [int(x) for x in a.pool_workers.split(",") if x]:
        pool = ActorPool([ActorPoolConfig(backend=a.pool_backend, count=k)])
        await pool.start()
        try:
            out[f"pool_x{k}_ms"] = round(await via(pool), 3)
        finally:
            await pool.shutdown()
    print(json.dumps(out))


if __name__ == "__main__":
    asyncio.run(main())
