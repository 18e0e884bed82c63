"""Single-process robust training loop: k simulated workers' gradients of one SmallCNN are aggregated
with Minimum-Diameter-Averaging through a NodeScheduler, directly and on an ActorPool (counterpart of
the reference's benchmarks/pytorch/mnist_training_actor_pool.py).

    python benchmarks/pytorch/mnist_training_actor_pool.py --rounds 3 --num-workers 14 --byz-workers 4 --f 4 \
        --chunk-size 128 --pool-workers 4 --pool-backend process

Flags and defaults are the reference script's; ``--steps`` / ``--workers`` / ``--byzantine`` are older aliases.
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from benchmarks.pytorch._worker_args import pool_configs  # noqa: E402
from byzpy_b200.aggregators.geometric_wise import MinimumDiameterAveraging  # noqa: E402
from byzpy_b200.attacks import SignFlipAttack  # noqa: E402
from byzpy_b200.engine.graph.ops import make_single_operator_graph  # noqa: E402
from byzpy_b200.engine.graph.pool import ActorPool  # noqa: E402
from byzpy_b200.engine.graph.scheduler import NodeScheduler  # noqa: E402
from byzpy_b200.models import SmallCNN  # noqa: E402
from byzpy_b200.parallel.arena import flatten_grads, write_vector_to_grads_  # noqa: E402
from byzpy_b200.utils.data import batch_source, evaluate, mnist_like, shard_indices  # noqa: E402


async def train(a, pool):
    steps, workers, n_byz = a.rounds, a.num_workers, a.byz_workers
    torch.manual_seed(a.seed)
    model = SmallCNN()
    opt = torch.optim.SGD(model.parameters(), lr=a.lr)
    lossf = torch.nn.CrossEntropyLoss()
    x, y = mnist_like(6000, root=a.data_root, seed=a.seed)
    srcs = [batch_source(x[torch.as_tensor(s)], y[torch.as_tensor(s)], a.batch_size, seed=a.seed + i)
            for i, s in enumerate(shard_indices(6000, workers))]
    graph = make_single_operator_graph(node_name="agg",
                                       operator=MinimumDiameterAveraging(f=a.f, chunk_size=a.chunk_size),
                                       input_keys=("gradients",))
    sched = NodeScheduler(graph, pool=pool)
    attack = SignFlipAttack()
    t_agg = 0.0
    for _ in range(steps):
        grads = []
        for w in range(workers):
            xb, yb = srcs[w]()
            model.zero_grad(set_to_none=True)
            lossf(model(xb), yb).backward()
            g = flatten_grads(model)
            grads.append(torch.as_tensor(attack.apply(base_grad=g)) if w >= workers - n_byz else g)
        t0 = time.perf_counter()
        agg = (await sched.run({"gradients": grads}))["agg"]
        t_agg += time.perf_counter() - t0
        write_vector_to_grads_(model, agg)
        opt.step()
    xt, yt = mnist_like(2000, train=False, root=a.data_root)
    loss, acc = evaluate(model, xt, yt, torch.device("cpu"))
    return t_agg / steps * 1e3, loss, acc


async def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-workers", "--workers", dest="num_workers", type=int, default=14)
    ap.add_argument("--byz-workers", "--byzantine", dest="byz_workers", type=int, default=4)
    ap.add_argument("--f", type=int, default=None, help="MDA f (default: --byz-workers)")
    ap.add_argument("--chunk-size", type=int, default=128, help="subsets scored per subtask")
    ap.add_argument("--rounds", "--steps", dest="rounds", type=int, default=3)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--lr", type=float, default=0.05)
    ap.add_argument("--pool-workers", type=int, default=4)
    ap.add_argument("--pool-backend", default="process")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--data-root", default="./data")
    a = ap.parse_args()
    a.f = a.byz_workers if a.f is None else a.f
    ms, loss, acc = await train(a, None)
    out = {"rounds": a.rounds, "workers": a.num_workers, "byzantine": a.byz_workers, "f": a.f,
           "direct_agg_ms": round(ms, 2), "direct_test_acc": round(acc, 4)}
    pool = ActorPool(pool_configs(a.pool_backend, a.pool_workers))
    await pool.start()
    try:
        ms, loss, acc = await train(a, pool)
        out[f"pool_x{a.pool_workers}_agg_ms"] = round(ms, 2)
        out["pool_test_acc"] = round(acc, 4)
    finally:
        await pool.shutdown()
    print(json.dumps(out))


if __name__ == "__main__":
    asyncio.run(main())
