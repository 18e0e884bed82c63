"""Device-timed throughput of the BASELINE.json training configs other than the headline
(`bench.py` is config 2):

  --config 3   ResNet-50 parameter server, 8 rows = 6 honest replicas + 2 Little rows (the omniscient
               adversary is two *virtual* rows synthesised inside the aggregation kernels from the honest
               rows' column mean / std), Bucketing -> Multi-Krum composed in n-space
  --config 4   BERT-base peer-to-peer (gossip), 8 peers = 7 honest + 1 Empire, GeometricMedian
               (Weiszfeld in Gram space), complete topology

    python benchmarks/training_configs.py --config 3 --steps 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        benchmarks/training_configs.py --config 4 --steps 10

Synthetic data of the named shapes, random-init weights, bf16 autocast, CUDA events around exactly
`--steps` rounds, max over ranks.  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from byzpy_b200.aggregators.geometric_wise import GeometricMedian, MultiKrum  # noqa: E402
from byzpy_b200.attacks import EmpireAttack, LittleAttack  # noqa: E402
from byzpy_b200.engine.node.device import (DeviceByzantineNode, DeviceHonestNode, DeviceP2PByzantineNode,  # noqa: E402
                                            DeviceP2PHonestNode)
from byzpy_b200.engine.parameter_server.ps import ParameterServer  # noqa: E402
from byzpy_b200.engine.peer_to_peer.topology import Topology  # noqa: E402
from byzpy_b200.engine.peer_to_peer.train import PeerToPeer  # noqa: E402
from byzpy_b200.models import build_model  # noqa: E402
from byzpy_b200.ops import normalize_uint8_nhwc  # noqa: E402
from byzpy_b200.parallel.device_p2p import PeerLayout  # noqa: E402
from byzpy_b200.parallel.device_ps import RowLayout  # noqa: E402
from byzpy_b200.pre_aggregators import Bucketing  # noqa: E402


def max_over_ranks(v: float, dev) -> float:
    if dist.is_initialized():
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return v


def sync(dev):
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize(dev)


def timed(step, steps, warmup, dev):
    for _ in range(max(3, warmup)):
        step()
    sync(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    sync(dev)
    return max_over_ranks(e0.elapsed_time(e1), dev)


def config3(a, rank, world, dev):
    n_h, n_virtual = 6, 2
    # 6 replicas over any number of ranks up to 8: ranks beyond the sixth host no replica and only take
    # part in the aggregation of their coordinate shard (RowLayout.spread)
    layout = RowLayout.spread(n_h, 0, world, n_virtual=n_virtual)
    g = torch.Generator().manual_seed(rank)
    pool = [(torch.randint(0, 256, (a.batch, 224, 224, 3), dtype=torch.uint8, generator=g).pin_memory(),
             torch.randint(0, 1000, (a.batch,), generator=g).pin_memory()) for _ in range(2)]
    cur = [0]

    def source():
        cur[0] += 1
        return pool[cur[0] % 2]

    honest = []
    for _ in layout.local_ids(rank):
        torch.manual_seed(0)
        honest.append(DeviceHonestNode(build_model("resnet50", num_classes=1000), lr=0.05, momentum=0.9,
                                       device=str(dev), data=source,
                                       preprocess=lambda x: normalize_uint8_nhwc(x, 127.5, 127.5, s2d=True)))
    # the two Little rows are virtual (synthesised in-kernel); every rank declares them
    byz = [DeviceByzantineNode(LittleAttack(f=n_virtual), device=str(dev)) for _ in range(n_virtual)]
    ps = ParameterServer(honest, byz, MultiKrum(f=1, q=2), pre_aggregator=Bucketing(bucket_size=2),
                         update_byzantines=False, layout=layout, fused=True, worker_streams=4, lr=0.05,
                         momentum=0.9)
    ms = timed(ps.step, a.steps, a.warmup, dev)
    ps.device_round.check_status()
    d = ps.device_round.d
    asyncio.run(ps.shutdown())
    return dict(metric="PS steps/sec (device-timed, max over ranks) ResNet-50 + Bucketing->Multi-Krum, 2 Little",
                value=round(a.steps / ms * 1e3, 3), ms_per_step=round(ms / a.steps, 3), grad_dim=d,
                config=dict(model="resnet50", rows="6 honest replicas + 2 virtual Little rows", batch=a.batch))


def config4(a, rank, world, dev):
    peers, n_b = 8, 1
    layout = PeerLayout(peers - n_b, n_b, world)
    g = torch.Generator().manual_seed(rank)
    ids = torch.randint(0, 30522, (a.batch, a.seq), generator=g).pin_memory()

    def source():
        return ids, ids

    loss = lambda out, y: torch.nn.functional.cross_entropy(out.flatten(0, 1), y.flatten())  # noqa: E731
    hon, byz = [], []
    for gid in layout.local_ids(rank):
        if gid < layout.n_honest:
            torch.manual_seed(0)
            hon.append(DeviceP2PHonestNode(build_model("bert-base"), GeometricMedian(), loss_fn=loss, data=source,
                                           device=str(dev)))
        else:
            byz.append(DeviceP2PByzantineNode(EmpireAttack(scale=-1.0), device=str(dev)))
    p2p = PeerToPeer(hon, byz, Topology.complete(peers), lr=0.01, layout=layout, fused=True,
                     amp_dtype=torch.bfloat16)
    ms = timed(p2p.step, a.steps, a.warmup, dev)
    p2p.device_round.check_status()
    asyncio.run(p2p.shutdown())
    return dict(metric="P2P rounds/sec (device-timed, max over ranks) BERT-base + GeometricMedian, 1 Empire",
                value=round(a.steps / ms * 1e3, 3), ms_per_step=round(ms / a.steps, 3),
                config=dict(model="bert-base", peers=peers, batch=a.batch, seq_len=a.seq))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, choices=[3, 4], required=True)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--seq", type=int, default=128)
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 32 if a.config == 3 else 8
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = True
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = (config3 if a.config == 3 else config4)(a, rank, world, dev)
    if rank == 0:
        out.update(unit="steps/s", n_gpus=world, steps=a.steps, dtype="bf16", data="synthetic")
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
