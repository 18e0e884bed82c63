"""Multi-rank correctness check of the fused cross-GPU parameter-server round.

Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
             --master-port 29533 tests/multi_gpu/check_fused_round.py [--agg median|trmean|meamed]

Every rank hosts 8/N TinyNet replicas; after each round the fused kernel's aggregate (delivered
into every rank's buffer by peer / multicast stores) is compared with an INDEPENDENT oracle: the
local gradients of mirror models are exchanged with an NCCL all_gather and aggregated on the host
in float64 with plain torch algebra (tests/multi_gpu/oracle.py -- none of this repo's operators),
and the updated parameters with a plain torch SGD loop.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import oracle  # noqa: E402

from byzpy_b200.aggregators.coordinate_wise import (CoordinateWiseMedian, CoordinateWiseTrimmedMean,  # noqa: E402
                                                     MeanOfMedians)
from byzpy_b200.aggregators.geometric_wise import GeometricMedian, MultiKrum  # noqa: E402
from byzpy_b200.aggregators.norm_wise import CenteredClipping  # noqa: E402
from byzpy_b200.attacks import LittleAttack, SignFlipAttack  # noqa: E402
from byzpy_b200.pre_aggregators import Bucketing  # noqa: E402
from byzpy_b200.engine.node.device import DeviceByzantineNode, DeviceHonestNode  # noqa: E402
from byzpy_b200.engine.parameter_server.ps import ParameterServer  # noqa: E402
from byzpy_b200.ops import reference as ref  # noqa: E402
from byzpy_b200.parallel.device_ps import RowLayout  # noqa: E402


class TinyNet(nn.Module):
    def __init__(self, width=257):
        super().__init__()
        self.a = nn.Linear(64, width)
        self.b = nn.Linear(width, width)
        self.c = nn.Linear(width, 10)

    def forward(self, x):
        return self.c(torch.relu(self.b(torch.relu(self.a(x)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agg", default="median")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--attack", default="signflip", choices=["signflip", "little"])
    ap.add_argument("--workers", type=int, default=8,
                    help="total rows; need not divide the world size (RowLayout.spread)")
    ap.add_argument("--buckets", type=int, default=0,
                    help="0 = engine default; k > 1 forces small gradient buckets so the overlapped "
                         "bucket protocol (per-bucket sequence numbers) is exercised across ranks")
    ap.add_argument("--multicast", type=int, default=-1, help="-1 auto, 0 peer stores, 1 require NVLS multicast")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_total, n_byz = a.workers, 2
    n_h = n_total - n_byz
    virtual = a.attack == "little"
    layout = (RowLayout.spread(n_h, 0, world, n_virtual=n_byz) if virtual
              else RowLayout.spread(n_h, n_byz, world))
    gids = layout.local_ids(rank)
    torch.manual_seed(0)
    init = TinyNet().state_dict()
    hon, byz, mirror = [], [], []
    for g in gids:
        m = TinyNet()
        m.load_state_dict(init)
        m2 = TinyNet()
        m2.load_state_dict(init)
        mirror.append(m2.to(dev))
        kw = dict(lr=0.1, momentum=0.9, device=str(dev))
        if g < n_h:
            hon.append(DeviceHonestNode(m, **kw))
        else:
            byz.append(DeviceByzantineNode(SignFlipAttack(), model=m, **kw))
    mk = {"median": lambda: CoordinateWiseMedian(), "trmean": lambda: CoordinateWiseTrimmedMean(f=2),
          "meamed": lambda: MeanOfMedians(f=2), "multikrum": lambda: MultiKrum(f=2, q=4),
          "gm": lambda: GeometricMedian(tol=1e-7), "gm_mean": lambda: GeometricMedian(init="mean"),
          "cclip": lambda: CenteredClipping(c_tau=0.5, M=8), "bucket_krum": lambda: MultiKrum(f=1, q=2)}[a.agg]
    agg = mk()
    pre = Bucketing(2, perm=[3, 0, 6, 1, 7, 2, 5, 4]) if a.agg == "bucket_krum" else None
    if virtual:
        byz = [DeviceByzantineNode(LittleAttack(f=n_byz), device=str(dev)) for _ in range(n_byz)]
    opts_dev = {}
    if a.buckets > 1:
        opts_dev = dict(min_bucket=2048, bucket_cuts=tuple((i + 1) / a.buckets for i in range(a.buckets - 1)))
    ps = ParameterServer(hon, byz, agg, pre_aggregator=pre, update_byzantines=True, layout=layout,
                         amp_dtype=None, use_cuda_graph=bool(a.graph), fused=True, lr=0.1, momentum=0.9,
                         multicast=None if a.multicast < 0 else bool(a.multicast), device_options=opts_dev)
    if a.buckets > 1 and a.agg in ("median", "trmean", "meamed"):
        assert ps.device_round.n_buckets > 1, ps.device_round._bounds
    opts = [torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9) for m in mirror]
    lossf = nn.CrossEntropyLoss()
    ok = True
    for t in range(a.steps):
        gen = torch.Generator().manual_seed(1000 * t + rank)
        batches = [(torch.randn(32, 64, generator=gen).pin_memory(),
                    torch.randint(0, 10, (32,), generator=gen).pin_memory()) for _ in gids]
        ps.step(batches)
        # reference: local grads on the mirrors, NCCL all_gather, torch aggregate, torch SGD
        local_rows = []
        for (x, y), m, g in zip(batches, mirror, gids):
            m.zero_grad()
            lossf(m(x.to(dev)), y.to(dev)).backward()
            v = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
            local_rows.append(-v if (g >= n_h and not virtual) else v)
        d_flat = sum(p.numel() for p in TinyNet().parameters())
        per = layout.max_local()
        loc = torch.zeros((per, d_flat), device=dev)
        for k, v in enumerate(local_rows):
            loc[k] = v
        full = torch.empty((world, per, d_flat), device=dev)
        dist.all_gather_into_tensor(full.view(-1), loc.view(-1))
        rows = [full[layout.rank_of[g], layout.slot_of[g]] for g in range(layout.n_workers)]
        if virtual:
            mal = oracle.little(rows, n_byz).to(dev, torch.float32)
            rows = rows + [mal] * n_byz
        expect = oracle.aggregate(a.agg, rows).to(dev, torch.float32)     # fp64 host algebra, not this repo's ops
        for m, o in zip(mirror, opts):
            off = 0
            for p in m.parameters():
                p.grad.copy_(expect[off:off + p.numel()].view_as(p))
                off += p.numel()
            o.step()
        torch.cuda.synchronize()
        ps.device_round.check_status()
        got = ps.device_round.aggregated()
        e1 = (got - expect).abs().max().item()
        e2 = 0.0
        if mirror:          # a rank without a replica only checks the delivered aggregate
            first = hon[0].model if hon else next(b.model for b in byz if b.model is not None)
            mine = torch.cat([p.detach().reshape(-1) for p in first.parameters()])
            theirs = torch.cat([p.detach().reshape(-1) for p in mirror[0].parameters()])
            e2 = (mine - theirs).abs().max().item()
        good = e1 < 2e-5 and e2 < 2e-4
        ok = ok and good
        print(f"[rank {rank}] step {t}: |agg-ref|={e1:.2e} |param-ref|={e2:.2e} {'OK' if good else 'MISMATCH'}",
              flush=True)
    if rank == 0:
        r = ps.device_round
        print(f"[rank 0] buckets={r.n_buckets} overlapped={r._use_buckets} multicast={bool(r._agg_mc)} "
              f"launches/round={r.launches_per_step}", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_FUSED_ROUND", "PASS" if flag.item() == 1.0 else "FAIL", f"world={world} agg={a.agg}", flush=True)
    import asyncio

    asyncio.run(ps.shutdown())
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
