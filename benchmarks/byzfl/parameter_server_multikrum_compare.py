"""Parameter-server training with Multi-Krum: this library's ``ParameterServer`` vs ByzFL's
``Server`` / ``Client`` / ``ByzantineClient`` loop on the same model, data shape, node counts and
round count (counterpart of the reference's benchmarks/byzfl/parameter_server_multikrum_compare.py,
the "ByzFL 57 ms" column of benchmarks/README.md:23).

Both arms train SmallCNN on the same in-memory MNIST-shaped tensors (real MNIST when ``--data-root``
holds a copy, otherwise ``utils.data.mnist_like``'s synthetic stand-in -- the build image has no network),
``--num-honest`` honest clients and ``--num-byz`` Byzantine ones (Empire here, SignFlipping in ByzFL), SGD with ``--lr``, for ``--rounds``
rounds; the report is total and per-round milliseconds.  ByzFL is not installable offline: its arm
then reads ``unavailable``.

    python benchmarks/byzfl/parameter_server_multikrum_compare.py --rounds 50 --num-honest 10 --num-byz 3 --f 3
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

from examples.ps.nodes import DistributedPSByzNode, DistributedPSHonestNode, SmallCNN  # noqa: E402

from byzpy_b200.aggregators.geometric_wise import MultiKrum  # noqa: E402
from byzpy_b200.engine.node.actors import ByzantineNodeActor, HonestNodeActor  # noqa: E402
from byzpy_b200.engine.parameter_server.ps import ParameterServer  # noqa: E402
from byzpy_b200.utils.data import mnist_like, shard_indices  # noqa: E402


async def ours(a) -> dict:
    shards = shard_indices(a.samples, a.honest)
    hon = [await HonestNodeActor.spawn(DistributedPSHonestNode, backend="thread",
                                       kwargs=dict(indices=shards[i], seed=a.seed + i, batch_size=a.batch_size,
                                                   lr=a.lr))
           for i in range(a.honest)]
    byz = [await ByzantineNodeActor.spawn(DistributedPSByzNode, backend="thread") for _ in range(a.byzantine)]
    ps = ParameterServer(hon, byz, MultiKrum(f=a.f, q=max(1, a.honest + a.byzantine - a.f - 1)))
    await ps.round()
    t0 = time.perf_counter()
    for _ in range(a.rounds):
        await ps.round()
    total = time.perf_counter() - t0
    await ps.shutdown()
    return {"total_ms": round(total * 1e3, 1), "ms_per_round": round(total / a.rounds * 1e3, 2)}


def theirs(a) -> dict:
    try:
        from byzfl import ByzantineClient, Client, Server
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": type(exc).__name__}
    x, y = mnist_like(a.samples, root=a.data_root, seed=a.seed)
    x = (x - 0.1307) / 0.3081
    shards = shard_indices(a.samples, a.honest)
    loaders = [torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x[s], y[s]), batch_size=a.batch_size,
                                           shuffle=True) for s in shards]
    common = {"model_name": None, "model": SmallCNN(), "device": "cpu", "loss_name": "CrossEntropyLoss",
              "learning_rate": a.lr, "weight_decay": 0.0, "milestones": [], "learning_rate_decay": 1.0,
              "LabelFlipping": False, "momentum": 0.0, "nb_labels": 10}
    clients = [Client({**common, "training_dataloader": ld}) for ld in loaders]
    server = Server({**common, "test_loader": None, "validation_loader": None,
                     "aggregator_info": {"name": "MultiKrum", "parameters": {"f": a.f}},
                     "pre_agg_list": []})
    attacker = ByzantineClient({"name": "SignFlipping", "f": a.byzantine, "parameters": {}})

    def one_round():
        for c in clients:
            c.compute_gradients()
        honest = [c.get_flat_gradients_with_momentum() for c in clients]
        server.update_model(honest + attacker.apply_attack(honest))
        state = server.get_dict_parameters()
        for c in clients:
            c.set_model_state(state)

    one_round()
    t0 = time.perf_counter()
    for _ in range(a.rounds):
        one_round()
    total = time.perf_counter() - t0
    return {"total_ms": round(total * 1e3, 1), "ms_per_round": round(total / a.rounds * 1e3, 2)}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=50)
    ap.add_argument("--num-honest", "--honest", dest="honest", type=int, default=10)
    ap.add_argument("--num-byz", "--byzantine", dest="byzantine", type=int, default=3)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--lr", type=float, default=0.1)
    ap.add_argument("--f", type=int, default=None, help="Multi-Krum f (default: --num-byz)")
    ap.add_argument("--samples", type=int, default=6000)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--data-root", default="./data")
    a = ap.parse_args()
    a.f = a.byzantine if a.f is None else a.f
    out = {"rounds": a.rounds, "honest": a.honest, "byzantine": a.byzantine, "batch_size": a.batch_size,
           "byzpy_b200": asyncio.run(ours(a)), "byzfl": theirs(a)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
