"""Fault-injection check of the fused parameter-server round (2 ranks).

Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
             --master-port 29539 tests/multi_gpu/check_fault.py

Rank 1 goes silent after two good rounds (``DeviceRound.inject_fault("silent")``).  Rank 0's fused
kernels time out on rank 1's flags (``spin_seconds``), ``read_losses()`` raises and names the silent
rank, ``ParameterServer.recover()`` drops rank 1's rows and rebuilds the plan, and training
continues on the remaining rows; every aggregate is compared with the independent fp64 oracle.
"""
import asyncio
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import oracle  # noqa: E402

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian  # noqa: E402
from byzpy_b200.engine.node.device import DeviceHonestNode  # noqa: E402
from byzpy_b200.engine.parameter_server.ps import ParameterServer  # noqa: E402
from byzpy_b200.parallel.device_ps import RowLayout  # noqa: E402


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(32, 64)
        self.b = nn.Linear(64, 10)

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    assert world == 2
    per = 3
    layout = RowLayout.block(2 * per, 0, world)
    torch.manual_seed(0)
    init = Net().state_dict()
    hon, mirror = [], []
    for _ in range(per):
        m = Net()
        m.load_state_dict(init)
        hon.append(DeviceHonestNode(m, lr=0.1, momentum=0.0, device=str(dev)))
        m2 = Net().to(dev)
        m2.load_state_dict(init)
        mirror.append(m2)
    ps = ParameterServer(hon, [], CoordinateWiseMedian(), layout=layout, amp_dtype=None, use_cuda_graph=True,
                         fused=True, lr=0.1, momentum=0.0, device_options=dict(spin_seconds=0.5))
    rnd = ps.device_round
    lossf = nn.CrossEntropyLoss()
    ok = True

    def local_rows(t):
        gen = torch.Generator().manual_seed(50 * t + rank)
        batches = [(torch.randn(16, 32, generator=gen).pin_memory(),
                    torch.randint(0, 10, (16,), generator=gen).pin_memory()) for _ in range(per)]
        rows = []
        for (x, y), m in zip(batches, mirror):
            m.zero_grad()
            lossf(m(x.to(dev)), y.to(dev)).backward()
            rows.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]))
        return batches, rows

    def apply(expect):
        with torch.no_grad():
            for m in mirror:
                off = 0
                for p in m.parameters():
                    p.add_(expect[off:off + p.numel()].view_as(p), alpha=-0.1)
                    off += p.numel()

    d = sum(p.numel() for p in Net().parameters())
    # ---- two healthy rounds
    for t in range(2):
        batches, rows = local_rows(t)
        ps.step(batches)
        full = torch.empty((world, per, d), device=dev)
        dist.all_gather_into_tensor(full.view(-1), torch.stack(rows).view(-1))
        expect = oracle.median(list(full.view(world * per, d).unbind(0))).to(dev, torch.float32)
        rnd.read_losses()
        err = (rnd.aggregated() - expect).abs().max().item()
        ok = ok and err < 2e-5
        apply(expect)
        print(f"[rank {rank}] healthy round {t}: |agg-ref|={err:.2e}", flush=True)
    dist.barrier()
    # ---- rank 1 goes silent
    if rank == 1:
        rnd.inject_fault("silent")
        ps.step([(torch.zeros(16, 32).pin_memory(), torch.zeros(16, dtype=torch.long).pin_memory())] * per)
        time.sleep(6.0)                      # stays alive (its memory stays mapped) but takes no part
        print("[rank 1] silent", flush=True)
    else:
        batches, rows = local_rows(2)
        ps.step(batches)
        raised = False
        try:
            rnd.read_losses()
        except RuntimeError as exc:
            raised = "silent ranks [1]" in str(exc)
            print(f"[rank 0] round failed as expected: {str(exc)[:160]}", flush=True)
        ok = ok and raised
        dropped = ps.recover()
        ok = ok and dropped == [1] and rnd.layout.n_workers == per and rnd.live_mask == 1
        print(f"[rank 0] recovered: dropped ranks {dropped}, rows left {rnd.layout.n_workers}", flush=True)
        for t in range(3, 6):               # training continues on the remaining rows
            batches, rows = local_rows(t)
            ps.step(batches)
            expect = oracle.median(rows).to(dev, torch.float32)
            rnd.read_losses()
            err = (rnd.aggregated() - expect).abs().max().item()
            apply(expect)
            mine = torch.cat([p.detach().reshape(-1) for p in hon[0].model.parameters()])
            theirs = torch.cat([p.detach().reshape(-1) for p in mirror[0].parameters()])
            perr = (mine - theirs).abs().max().item()
            ok = ok and err < 2e-5 and perr < 2e-4
            print(f"[rank 0] degraded round {t}: |agg-ref|={err:.2e} |param-ref|={perr:.2e}", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_FAULT", "PASS" if flag.item() == 1.0 else "FAIL", flush=True)
    asyncio.run(ps.shutdown())
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
