"""Parameter-server training with node actors on the process backend (each node in its own OS process) (generic actor path):
4 honest SmallCNN workers + 1 Empire Byzantine worker, CoordinateWiseMedian.

    python examples/ps/thread/mnist.py [--rounds 100] [--backend thread|process]
"""
from __future__ import annotations

import argparse
import asyncio
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from examples.ps.nodes import DistributedPSByzNode, DistributedPSHonestNode, select_pool_backend  # noqa: E402

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian  # noqa: E402
from byzpy_b200.configs.actor import set_actor  # noqa: E402
from byzpy_b200.engine.node.actors import ByzantineNodeActor, HonestNodeActor  # noqa: E402
from byzpy_b200.engine.parameter_server.ps import ParameterServer  # noqa: E402
from byzpy_b200.models import SmallCNN  # noqa: E402
from byzpy_b200.utils import train_with_progress  # noqa: E402
from byzpy_b200.utils.data import evaluate, mnist_like, shard_indices  # noqa: E402


async def main(rounds: int, backend: str):
    n_honest, n_byz = 4, 1
    shards = shard_indices(6000, n_honest)
    pool_backend = select_pool_backend(backend)
    honest = [await HonestNodeActor.spawn(DistributedPSHonestNode, backend=set_actor(backend),
                                          kwargs=dict(indices=shards[i], pool_backend=pool_backend, seed=i))
              for i in range(n_honest)]
    byz = [await ByzantineNodeActor.spawn(DistributedPSByzNode, backend=set_actor(backend),
                                          kwargs=dict(pool_backend=pool_backend)) for _ in range(n_byz)]
    ps = ParameterServer(honest, byz, CoordinateWiseMedian(), update_byzantines=False)
    xt, yt = mnist_like(2000, train=False)
    probe = SmallCNN()

    async def evaluate_now():
        probe.load_state_dict(await honest[0].dump_state_dict(), strict=True)
        loss, acc = evaluate(probe, xt, yt, torch.device("cpu"))
        return {"loss": round(loss, 4), "acc": round(acc, 4)}

    await train_with_progress(ps, rounds, eval_callback=evaluate_now, eval_interval=max(1, rounds // 4))
    print("final:", await evaluate_now())
    await ps.shutdown()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=100)
    ap.add_argument("--backend", default="process")
    a = ap.parse_args()
    asyncio.run(main(a.rounds, a.backend))
