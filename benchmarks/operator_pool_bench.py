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


# Per-operator defaults of the reference's scripts (benchmarks/pytorch/<op>_actor_pool.py / *_preagg.py): the
# same-named front-ends here accept the same flags with the same defaults, so a command line written for the
# reference keeps working (`multikrum_actor_pool.py --num-grads 80 --f 20 --q 12 --chunk-size 20 ...`).
REF_DEFAULTS = {
    "median": dict(n=64, d=65536, chunk=8192),
    "trimmed-mean": dict(n=64, d=65536, f=8, chunk=8192),
    "meamed": dict(n=64, d=65536, f=8, chunk=8192),
    "multi-krum": dict(n=64, d=65536, f=8, q=8, chunk=16),
    "krum": dict(n=64, d=65536, f=8, chunk=16),
    "geometric-median": dict(n=64, d=65536, chunk=16),
    "mda": dict(n=18, d=2048, f=6, chunk=256),
    "monna": dict(n=64, d=65536, f=8, chunk=32),
    "smea": dict(n=12, d=1024, f=3, chunk=128),
    "centered-clipping": dict(n=64, d=131072, chunk=16),
    "cge": dict(n=128, d=131072, f=16, chunk=32768),
    "caf": dict(n=64, d=65536, f=8, chunk=32),
    "clipping": dict(n=256, d=65536, chunk=32),
    "arc": dict(n=256, d=65536, f=8, chunk=32),
    "nnm": dict(n=256, d=65536, f=32, chunk=16384),
    "bucketing": dict(n=512, d=16384, chunk=8192),
    "empire": dict(n=128, d=131072, chunk=32),
    "little": dict(n=128, d=131072, f=16, chunk=16384),
    "gaussian": dict(n=64, d=65536, chunk=16384),
    "inf": dict(n=64, d=65536, chunk=16384),
    "mimic": dict(n=64, d=65536, chunk=16384),
    "sign-flip": dict(n=1, d=262144, chunk=8192),
}


def make(op: str, n: int, f: int, a=None):
    """(factory, input key) of operator ``op``; ``a`` = parsed arguments carrying the reference's per-script
    knobs (all optional)."""
    g = lambda name, default: default if a is None or getattr(a, name, None) is None else getattr(a, name)  # noqa: E731
    chunk = g("chunk_size", None)
    ck = {} if chunk is None else {"chunk_size": int(chunk)}
    table = {
        "median": (lambda: CoordinateWiseMedian(**ck), "gradients"),
        "trimmed-mean": (lambda: CoordinateWiseTrimmedMean(f=f, **ck), "gradients"),
        "meamed": (lambda: MeanOfMedians(f=f, **ck), "gradients"),
        "multi-krum": (lambda: MultiKrum(f=f, q=int(g("q", max(1, n - 2 * f))), **ck), "gradients"),
        "krum": (lambda: Krum(f=f, **ck), "gradients"),
        "geometric-median": (lambda: GeometricMedian(tol=float(g("tol", 1e-6)), max_iter=int(g("max_iter", 128)),
                                                     init=g("init", "median"), **ck), "gradients"),
        "mda": (lambda: MinimumDiameterAveraging(f=f, **ck), "gradients"),
        "monna": (lambda: MoNNA(f=f, reference_index=int(g("reference_index", 0)), **ck), "gradients"),
        "smea": (lambda: SMEA(f=f, **ck), "gradients"),
        "centered-clipping": (lambda: CenteredClipping(c_tau=float(g("c_tau", 0.1)), M=int(g("iters", 10)),
                                                      **({} if g("init", None) is None else {"init": g("init", "mean")}),
                                                      **ck), "gradients"),
        "cge": (lambda: ComparativeGradientElimination(f=f, **ck), "gradients"),
        "caf": (lambda: CAF(f=f, power_iters=int(g("power_iters", 3)), **ck), "gradients"),
        "clipping": (lambda: Clipping(threshold=float(g("threshold", 2.0)), **ck), "vectors"),
        "arc": (lambda: ARC(f=f, **ck), "vectors"),
        "nnm": (lambda: NearestNeighborMixing(f=f, **({} if (g("feature_chunk", None) or chunk) is None else
                                                      {"feature_chunk_size": int(g("feature_chunk", None) or chunk)})),
                "vectors"),
        "bucketing": (lambda: Bucketing(bucket_size=int(g("bucket_size", max(1, n // 16))),
                                        **({} if g("feature_chunk", None) is None
                                           else {"feature_chunk_size": int(g("feature_chunk", 8192))})), "vectors"),
        "empire": (lambda: EmpireAttack(scale=float(g("scale", -1.0)), **ck), "honest_grads"),
        "little": (lambda: LittleAttack(f=f, N=g("N", None), **ck), "honest_grads"),
        "gaussian": (lambda: GaussianAttack(mu=float(g("mu", 0.0)), sigma=float(g("sigma", 1.0)), seed=0, **ck),
                     "honest_grads"),
        "inf": (lambda: InfAttack(**ck), "honest_grads"),
        "mimic": (lambda: MimicAttack(epsilon=int(g("epsilon", 0)), **ck), "honest_grads"),
        "sign-flip": (lambda: SignFlipAttack(scale=float(g("scale", -1.0)), **ck), "base_grad"),
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
    # sizes: None = the reference script's default for --op (REF_DEFAULTS); --num-vectors / --dim are the names the
    # reference's pre-aggregator scripts use
    ap.add_argument("--num-grads", "--num-vectors", dest="num_grads", type=int, default=None)
    ap.add_argument("--grad-dim", "--dim", dest="grad_dim", type=int, default=None)
    ap.add_argument("--f", type=int, default=None)
    ap.add_argument("--chunk-size", type=int, default=None, help="the operator's subtask granularity")
    for flag, typ in (("--q", int), ("--bucket-size", int), ("--feature-chunk", int), ("--threshold", float),
                      ("--c-tau", float), ("--iters", int), ("--tol", float), ("--max-iter", int), ("--mu", float),
                      ("--sigma", float), ("--scale", float), ("--epsilon", int), ("--N", int),
                      ("--reference-index", int), ("--power-iters", int)):
        ap.add_argument(flag, type=typ, default=None, help="operator knob of the reference script of the same name")
    ap.add_argument("--init", choices=["median", "mean"], default=None)
    ap.add_argument("--pool-workers", default="2,4,6")
    ap.add_argument("--pool-backend", default="thread")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    ref = REF_DEFAULTS.get(a.op, {})
    if a.num_grads is None:
        a.num_grads = ref.get("n", 64)
    if a.grad_dim is None:
        a.grad_dim = ref.get("d", 65536)
    explicit_f = a.f is not None
    if a.f is None:
        a.f = ref.get("f", 8)
    if a.chunk_size is None and "chunk" in ref:
        a.chunk_size = ref["chunk"]
    if a.q is None and "q" in ref:
        a.q = ref["q"]
    dev = torch.device(a.device)
    g = torch.Generator().manual_seed(a.seed)
    data = [torch.randn(a.grad_dim, generator=g).to(dev) for _ in range(a.num_grads)]
    # an explicit --f is taken as given (the operator validates it); a default is clamped to what n allows
    f = a.f if explicit_f else min(a.f, max(0, (a.num_grads - 1) // 2 - 1))
    mk, key = make(a.op, a.num_grads, f, a)
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
