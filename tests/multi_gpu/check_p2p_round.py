"""Multi-rank correctness check of the device peer-to-peer (gossip) round.

Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
             --master-port 29534 tests/multi_gpu/check_p2p_round.py [--agg trmean|gm] [--topology complete|ring]

8 peers (7 honest TinyNets + 1 Empire) are block-distributed over the ranks.  After every round each
local honest peer's parameters are compared with an INDEPENDENT oracle: the mirrors' half-step
vectors are exchanged with an NCCL all_gather and the Empire vectors / robust aggregates are computed
on the host in float64 with plain torch algebra (tests/multi_gpu/oracle.py -- none of this repo's
operators), following the P2P mixin semantics.
"""
import argparse
import asyncio
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import oracle  # noqa: E402

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseTrimmedMean  # noqa: E402
from byzpy_b200.aggregators.geometric_wise import GeometricMedian  # noqa: E402
from byzpy_b200.attacks import EmpireAttack  # noqa: E402
from byzpy_b200.engine.node.device import DeviceP2PByzantineNode, DeviceP2PHonestNode  # noqa: E402
from byzpy_b200.engine.peer_to_peer.topology import Topology  # noqa: E402
from byzpy_b200.engine.peer_to_peer.train import PeerToPeer  # noqa: E402
from byzpy_b200.parallel.device_p2p import PeerLayout  # noqa: E402


class TinyNet(nn.Module):
    def __init__(self, width=129):
        super().__init__()
        self.a = nn.Linear(32, width)
        self.b = nn.Linear(width, 10)

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def flat(m):
    return torch.cat([p.detach().reshape(-1) for p in m.parameters()])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agg", default="trmean", choices=["trmean", "gm"])
    ap.add_argument("--topology", default="complete", choices=["complete", "ring"])
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    peers, n_b, lr = 8, 1, 0.1
    n_h = peers - n_b
    layout = PeerLayout(n_h, n_b, world)
    gids = layout.local_ids(rank)
    topo = Topology.complete(peers) if a.topology == "complete" else Topology.ring(peers, 2)
    mk_agg = (lambda: CoordinateWiseTrimmedMean(f=1)) if a.agg == "trmean" else (lambda: GeometricMedian(tol=1e-7))
    torch.manual_seed(0)
    init = TinyNet().state_dict()
    hon, byz, mirror, honest_gids = [], [], {}, []
    for g in gids:
        if g < n_h:
            m = TinyNet()
            m.load_state_dict(init)
            hon.append(DeviceP2PHonestNode(m, mk_agg(), device=str(dev)))
            m2 = TinyNet().to(dev)
            m2.load_state_dict(init)
            mirror[g] = m2
            honest_gids.append(g)
        else:
            byz.append(DeviceP2PByzantineNode(EmpireAttack(scale=-2.0), device=str(dev)))
    p2p = PeerToPeer(hon, byz, topo, lr=lr, layout=layout, fused=True, amp_dtype=None, use_cuda_graph=True)
    lossf = nn.CrossEntropyLoss()
    d = sum(p.numel() for p in TinyNet().parameters())
    per = peers // world
    ok = True
    for t in range(a.steps):
        gen = torch.Generator().manual_seed(100 * t + rank)
        batches = []
        for g in gids:
            batches.append((torch.randn(16, 32, generator=gen).pin_memory(),
                            torch.randint(0, 10, (16,), generator=gen).pin_memory()) if g < n_h else None)
        p2p.step(batches)
        # reference: half steps on the mirrors
        halves = torch.zeros((per, d), device=dev)
        for slot, g in enumerate(gids):
            if g < n_h:
                m = mirror[g]
                x, y = batches[slot]
                m.zero_grad()
                lossf(m(x.to(dev)), y.to(dev)).backward()
                with torch.no_grad():
                    for p in m.parameters():
                        p.add_(p.grad, alpha=-lr)
                halves[slot] = flat(m)
        full = torch.empty((world, per, d), device=dev)
        dist.all_gather_into_tensor(full.view(-1), halves.view(-1))
        vec = {g: full[g // per, g % per] for g in range(peers)}
        for g in range(n_h, peers):       # Empire peers: attack on their honest in-neighbours
            ins = [j for j in dict.fromkeys(topo.in_.get(g, [])) if j < n_h]
            vec[g] = oracle.empire([vec[j] for j in ins], scale=-2.0).to(dev, torch.float32)
        torch.cuda.synchronize()
        p2p.device_round.check_status()
        for slot, g in enumerate(gids):
            if g >= n_h:
                continue
            ins = [j for j in dict.fromkeys(topo.in_.get(g, [])) if j != g]
            rows_g = [vec[g]] + [vec[j] for j in ins]
            expect = (oracle.trimmed_mean(rows_g, 1) if a.agg == "trmean"
                      else oracle.geometric_median(rows_g, tol=1e-7)).to(dev, torch.float32)
            got = p2p.device_round.param_vector(slot)
            err = (got - expect).abs().max().item()
            good = err < 2e-4
            ok = ok and good
            with torch.no_grad():           # the mirror continues from the aggregate, like the real peer
                off = 0
                for p in mirror[g].parameters():
                    p.copy_(expect[off:off + p.numel()].view_as(p))
                    off += p.numel()
            print(f"[rank {rank}] step {t} peer {g}: |theta-ref|={err:.2e} {'OK' if good else 'MISMATCH'}", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_P2P_ROUND", "PASS" if flag.item() == 1.0 else "FAIL",
              f"world={world} agg={a.agg} topology={a.topology}", flush=True)
    asyncio.run(p2p.shutdown())
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
