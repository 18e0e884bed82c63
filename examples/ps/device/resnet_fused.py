"""The B200-native path: device nodes + fused parameter-server round.

Single GPU:   python examples/ps/device/resnet_fused.py --rounds 20
Multi GPU:    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
                  examples/ps/device/resnet_fused.py --rounds 20

8 workers (6 honest + 2 SignFlip) are block-distributed over the ranks; every round is one CUDA-graph
launch per rank: fwd/bwd of the local replicas + ONE fused kernel (P2P gather, median, P2P broadcast,
SGD).  ``--checkpoint`` saves / resumes the full training state.
"""
from __future__ import annotations

import argparse
import asyncio
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian  # noqa: E402
from byzpy_b200.attacks import SignFlipAttack  # noqa: E402
from byzpy_b200.engine.node.device import DeviceByzantineNode, DeviceHonestNode  # noqa: E402
from byzpy_b200.engine.parameter_server.ps import ParameterServer  # noqa: E402
from byzpy_b200.models import build_model  # noqa: E402
from byzpy_b200.parallel.device_ps import RowLayout  # noqa: E402
from byzpy_b200.utils.checkpoint import load_checkpoint, save_checkpoint  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--model", default="resnet18")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--image", type=int, default=64)
    ap.add_argument("--classes", type=int, default=100)
    ap.add_argument("--checkpoint", default=None)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    layout = RowLayout.block(6, 2, world)
    gids = layout.local_ids(rank)
    g = torch.Generator().manual_seed(rank)

    def source():
        return (torch.randn(a.batch, 3, a.image, a.image, generator=g).pin_memory(),
                torch.randint(0, a.classes, (a.batch,), generator=g).pin_memory())

    honest, byz = [], []
    for gid in gids:
        torch.manual_seed(0)
        model = build_model(a.model, num_classes=a.classes)
        kw = dict(data=source, lr=0.05, momentum=0.9, device=str(dev))
        if gid < 6:
            honest.append(DeviceHonestNode(model, **kw))
        else:
            byz.append(DeviceByzantineNode(SignFlipAttack(), model=model, **kw))
    ps = ParameterServer(honest, byz, CoordinateWiseMedian(), update_byzantines=True, layout=layout, fused=True)
    start = 0
    if a.checkpoint and os.path.exists(a.checkpoint + f".rank{rank}"):
        start = load_checkpoint(a.checkpoint + f".rank{rank}", ps)
        print(f"[rank {rank}] resumed at round {start}")
    for r in range(start, a.rounds):
        ps.step()
        if rank == 0 and (r + 1) % 5 == 0:
            print(f"round {r + 1}: losses {ps.device_round.read_losses().tolist()}")
    ps.device_round.check_status()
    if a.checkpoint:
        save_checkpoint(a.checkpoint + f".rank{rank}", ps, round_index=a.rounds)
    asyncio.run(ps.shutdown())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
