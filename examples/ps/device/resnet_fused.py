"""The B200-native path: device nodes + fused parameter-server round.

Single GPU:   python examples/ps/device/resnet_fused.py --rounds 20
Multi GPU:    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
                  examples/ps/device/resnet_fused.py --rounds 20

8 workers (6 honest + 2 SignFlip) are block-distributed over the ranks; every round is one CUDA-graph
launch per rank: fwd/bwd of the local replicas + the fused kernel per gradient bucket (P2P gather, robust
aggregate, NVLS multicast / P2P broadcast, SGD), the late layers' buckets overlapped with backward.

  --aggregator median|trmean|multikrum|geomed   --pre none|bucketing|nnm      (any combination is one fused plan)
  --buckets K        gradient buckets (default: automatic; 1 = one launch after backward)
  --no-multicast     deliver with peer stores instead of multimem.st
  --timeline         print the device-side timeline of the last round (bucket launches vs backward)
  --silence-rank R --silence-at T   failure rehearsal: rank R stops taking part at round T; the others time out
                     (--spin-seconds), drop its rows (ParameterServer.recover) and continue
  --checkpoint PATH  save / resume the full training state
"""
from __future__ import annotations

import argparse
import asyncio
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean  # noqa: E402
from byzpy_b200.aggregators.geometric_wise import GeometricMedian, MultiKrum  # noqa: E402
from byzpy_b200.attacks import SignFlipAttack  # noqa: E402
from byzpy_b200.pre_aggregators import Bucketing, NearestNeighborMixing  # noqa: E402
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
    ap.add_argument("--aggregator", default="median", choices=["median", "trmean", "multikrum", "geomed"])
    ap.add_argument("--pre", default="none", choices=["none", "bucketing", "nnm"])
    ap.add_argument("--buckets", type=int, default=None)
    ap.add_argument("--no-multicast", action="store_true")
    ap.add_argument("--timeline", action="store_true")
    ap.add_argument("--silence-rank", type=int, default=-1)
    ap.add_argument("--silence-at", type=int, default=5)
    ap.add_argument("--spin-seconds", type=float, default=0.0, help="budget of device-side waits (0 = 20 s)")
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
    agg = {"median": lambda: CoordinateWiseMedian(), "trmean": lambda: CoordinateWiseTrimmedMean(f=2),
           "multikrum": lambda: MultiKrum(f=2, q=4), "geomed": lambda: GeometricMedian()}[a.aggregator]()
    pre = {"none": lambda: None, "bucketing": lambda: Bucketing(bucket_size=2),
           "nnm": lambda: NearestNeighborMixing(f=2)}[a.pre]()
    if pre is not None and a.aggregator in ("trmean", "multikrum") and a.pre == "bucketing":
        agg = {"trmean": CoordinateWiseTrimmedMean(f=1), "multikrum": MultiKrum(f=1, q=2)}[a.aggregator]   # 4 buckets
    if pre is not None and a.aggregator == "geomed":
        agg = GeometricMedian(init="mean")      # a median start row is not a linear function of the mixed rows
    opts = {}
    if a.spin_seconds or a.silence_rank >= 0:
        opts["spin_seconds"] = a.spin_seconds or 5.0
    if a.timeline:
        opts["trace"] = True
    ps = ParameterServer(honest, byz, agg, pre_aggregator=pre, update_byzantines=True, layout=layout, fused=True,
                         buckets=a.buckets, multicast=False if a.no_multicast else None, device_options=opts)
    rnd = ps.device_round
    if rank == 0:
        print(f"plan {type(rnd.plan).__name__}, {rnd.n_buckets} bucket(s), heap {rnd.sym.kind}, "
              f"multicast {bool(rnd.sym.mc_base)}, cuda graph {rnd.use_cuda_graph}")
    start = 0
    if a.checkpoint and os.path.exists(a.checkpoint + f".rank{rank}"):
        start = load_checkpoint(a.checkpoint + f".rank{rank}", ps)
        print(f"[rank {rank}] resumed at round {start}")
    for r in range(start, a.rounds):
        if rank == a.silence_rank and r == a.silence_at:
            rnd.inject_fault("silent")                 # this rank stops taking part, like a hung process
            print(f"[rank {rank}] going silent at round {r}")
        ps.step()
        if (r + 1) % 5 == 0 or (a.silence_rank >= 0 and r >= a.silence_at):
            try:
                losses = rnd.read_losses().tolist()
                if rank == 0:
                    print(f"round {r + 1}: losses {losses}")
            except RuntimeError as exc:                # a peer did not arrive within the wait budget
                if rank == a.silence_rank:
                    continue
                print(f"[rank {rank}] round {r + 1}: {str(exc)[:90]}...")
                dropped = ps.recover()
                print(f"[rank {rank}] dropped ranks {dropped}; continuing with {rnd.layout.n_workers} rows")
    if rank != a.silence_rank:
        rnd.check_status()
    if a.timeline and rank == 0:
        import json

        print(json.dumps(rnd.timeline(), indent=1))
    if a.checkpoint:
        save_checkpoint(a.checkpoint + f".rank{rank}", ps, round_index=a.rounds)
    asyncio.run(ps.shutdown())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
