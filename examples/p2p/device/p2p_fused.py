"""B200-native gossip training: device P2P nodes, vectors published in CUDA-IPC symmetric memory,
robust aggregation straight from peer HBM (BASELINE config 4 uses --model bert-base --agg gm).

    python examples/p2p/device/p2p_fused.py --rounds 10 --model smallcnn
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/p2p/device/p2p_fused.py --model bert-base --agg gm
"""
from __future__ import annotations

import argparse
import asyncio
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseTrimmedMean  # noqa: E402
from byzpy_b200.aggregators.geometric_wise import GeometricMedian  # noqa: E402
from byzpy_b200.attacks import EmpireAttack  # noqa: E402
from byzpy_b200.engine.node.device import DeviceP2PByzantineNode, DeviceP2PHonestNode  # noqa: E402
from byzpy_b200.engine.peer_to_peer.topology import Topology  # noqa: E402
from byzpy_b200.engine.peer_to_peer.train import PeerToPeer  # noqa: E402
from byzpy_b200.models import build_model  # noqa: E402
from byzpy_b200.parallel.device_p2p import PeerLayout  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--model", default="smallcnn")
    ap.add_argument("--agg", default="trmean", choices=["trmean", "gm"])
    ap.add_argument("--peers", type=int, default=8)
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_b = 1
    layout = PeerLayout(a.peers - n_b, n_b, world)
    gids = layout.local_ids(rank)
    g = torch.Generator().manual_seed(rank)
    bert = a.model.startswith("bert")

    def source():
        if bert:
            ids = torch.randint(0, 30522, (8, 128), generator=g)
            return ids.pin_memory(), ids.clone().pin_memory()
        return torch.rand(64, 1, 28, 28, generator=g).pin_memory(), torch.randint(0, 10, (64,), generator=g).pin_memory()

    loss = (lambda out, y: torch.nn.functional.cross_entropy(out.flatten(0, 1), y.flatten())) if bert else None
    hon, byz = [], []
    for gid in gids:
        if gid < layout.n_honest:
            torch.manual_seed(0)
            agg = GeometricMedian() if a.agg == "gm" else CoordinateWiseTrimmedMean(f=1)
            hon.append(DeviceP2PHonestNode(build_model(a.model), agg, loss_fn=loss, data=source, device=str(dev)))
        else:
            byz.append(DeviceP2PByzantineNode(EmpireAttack(scale=-1.0), device=str(dev)))
    p2p = PeerToPeer(hon, byz, Topology.complete(a.peers), lr=0.05, layout=layout, fused=True,
                     amp_dtype=torch.bfloat16)
    for r in range(a.rounds):
        p2p.step()
        if rank == 0 and (r + 1) % 5 == 0:
            print(f"round {r + 1}: losses {p2p.device_round.read_losses().tolist()}")
    p2p.device_round.check_status()
    asyncio.run(p2p.shutdown())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
