"""Single-process robust training loop: k simulated workers' gradients of one SmallCNN are aggregated
with Minimum-Diameter-Averaging through a NodeScheduler, directly and on an ActorPool (counterpart of
the reference's benchmarks/pytorch/mnist_training_actor_pool.py).

    python benchmarks/pytorch/mnist_training_actor_pool.py --steps 20 --workers 8 --byzantine 2 --pool-workers 4
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


async def train(steps, workers, n_byz, pool):
    torch.manual_seed(0)
    model = SmallCNN()
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    lossf = torch.nn.CrossEntropyLoss()
    x, y = mnist_like(6000)
    srcs = [batch_source(x[torch.as_tensor(s)], y[torch.as_tensor(s)], 32, seed=i)
            for i, s in enumerate(shard_indices(6000, workers))]
    graph = make_single_operator_graph(node_name="agg", operator=MinimumDiameterAveraging(f=n_byz),
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
    xt, yt = mnist_like(2000, train=False)
    loss, acc = evaluate(model, xt, yt, torch.device("cpu"))
    return t_agg / steps * 1e3, loss, acc


async def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--byzantine", type=int, default=2)
    ap.add_argument("--pool-workers", type=int, default=4)
    ap.add_argument("--pool-backend", default="thread")
    a = ap.parse_args()
    ms, loss, acc = await train(a.steps, a.workers, a.byzantine, None)
    out = {"steps": a.steps, "workers": a.workers, "byzantine": a.byzantine,
           "direct_agg_ms": round(ms, 2), "direct_test_acc": round(acc, 4)}
    pool = ActorPool(pool_configs(a.pool_backend, a.pool_workers))
    await pool.start()
    try:
        ms, loss, acc = await train(a.steps, a.workers, a.byzantine, pool)
        out[f"pool_x{a.pool_workers}_agg_ms"] = round(ms, 2)
        out["pool_test_acc"] = round(acc, 4)
    finally:
        await pool.shutdown()
    print(json.dumps(out))


if __name__ == "__main__":
    asyncio.run(main())
