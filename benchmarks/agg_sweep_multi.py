"""Multi-GPU aggregator sweep (BASELINE.json config 5): 8 gradient rows of d floats spread over N ranks
(8/N rows each, resident in CUDA-IPC symmetric memory), the fused cross-GPU aggregation round of this
library -- P2P gather over NVLink + aggregator + P2P broadcast + SGD on the local replicas, no NCCL --
against the reference's data path: NCCL all_gather of the rows + the reference aggregator math in torch
+ local SGD.  Device-timed (CUDA events, max over ranks), one JSON line per (aggregator, d).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        benchmarks/agg_sweep_multi.py --dims 1e6,1e7,1e8 --aggs median,trmean,multikrum,cclip
"""
from __future__ import annotations

import argparse
import asyncio
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean  # noqa: E402
from byzpy_b200.aggregators.geometric_wise import MultiKrum  # noqa: E402
from byzpy_b200.aggregators.norm_wise import CenteredClipping  # noqa: E402
from byzpy_b200.engine.node.device import DeviceHonestNode  # noqa: E402
from byzpy_b200.engine.parameter_server.ps import ParameterServer  # noqa: E402
from byzpy_b200.parallel.device_ps import RowLayout  # noqa: E402

N_ROWS = 8


class Flat(nn.Module):
    """A 'model' that is just one flat parameter vector: its gradient row is what gets aggregated."""

    def __init__(self, d: int):
        super().__init__()
        self.p = nn.Parameter(torch.zeros(d))

    def forward(self, x):
        return (self.p * x).sum().reshape(1, 1)


def make_agg(name: str):
    return {"median": lambda: CoordinateWiseMedian(), "trmean": lambda: CoordinateWiseTrimmedMean(f=2),
            "multikrum": lambda: MultiKrum(f=2, q=4), "cclip": lambda: CenteredClipping(c_tau=10.0, M=10)}[name]()


def ref_math(name: str, X: torch.Tensor) -> torch.Tensor:
    if name == "median":
        return X.median(dim=0).values
    if name == "trmean":
        return X.sort(dim=0).values[2:-2].mean(0)
    if name == "multikrum":
        G = X @ X.T
        dg = G.diag()
        D = (dg[:, None] + dg[None, :] - 2 * G).clamp_min(0)
        sc = D.sort(dim=1).values[:, 1:N_ROWS - 2].sum(1)
        return X[sc.argsort()[:4]].mean(0)
    v = X.median(dim=0).values
    for _ in range(10):
        diff = X - v
        a = (10.0 / diff.norm(dim=1).clamp_min(1e-12)).clamp(max=1.0)
        v = v + (a[:, None] * diff).mean(0)
    return v


def timeit(fn, dev, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="1e6,1e7,1e8")
    ap.add_argument("--aggs", default="median,trmean,multikrum,cclip")
    ap.add_argument("--skip-ref-above", type=float, default=2e8)
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = N_ROWS // world
    for d in [int(float(x)) for x in a.dims.split(",")]:
        for name in a.aggs.split(","):
            layout = RowLayout.block(N_ROWS, 0, world)
            nodes = [DeviceHonestNode(Flat(d), lr=0.1, momentum=0.9, device=str(dev)) for _ in range(L)]
            ps = ParameterServer(nodes, [], make_agg(name), layout=layout, fused=True, amp_dtype=None,
                                 use_cuda_graph=False, direct_grads=False)
            rnd = ps.device_round
            g = torch.Generator(device=dev).manual_seed(1000 * rank + 7)
            for w in rnd.workers:
                w.arena.flat_grads[:d].copy_(torch.randn(d, device=dev, generator=g))
            ours = timeit(rnd.launch_aggregate, dev)
            rnd.check_status()
            # bytes one rank moves per round: its coordinate shard of all 8 rows (once for coordinate-wise
            # aggregators, twice -- Gram pass + weighted-sum pass -- for the Gram family), the aggregate
            # written into every rank's buffer, and the SGD(+momentum) update of its L local replicas
            # (read p, m; write p, m -- the aggregate stays in registers)
            passes = 1 if name in ("median", "trmean") else 2
            per_gpu = (passes * N_ROWS * d / world + d + L * 4 * d) * 4
            out = dict(agg=name, d=d, n_gpus=world, rows=N_ROWS, ours_ms=round(ours, 4),
                       gathered_GB=round(N_ROWS * d * 4 / 1e9, 3),
                       bytes_per_gpu_GB=round(per_gpu / 1e9, 3),
                       ours_GBps_per_gpu=round(per_gpu / ours / 1e6, 1),
                       frac_of_hbm_peak=round(per_gpu / ours / 1e6 / 6482.7, 3))
            if d <= a.skip_ref_above:
                rows = torch.stack([w.arena.flat_grads[:d] for w in rnd.workers])
                params = [torch.zeros(d, device=dev) for _ in range(L)]
                moms = [torch.zeros(d, device=dev) for _ in range(L)]
                gathered = torch.empty(N_ROWS, d, device=dev)

                def reference():
                    if world > 1:
                        dist.all_gather_into_tensor(gathered.view(-1), rows.view(-1))
                    else:
                        gathered.copy_(rows)
                    agg = ref_math(name, gathered)
                    for p, m in zip(params, moms):
                        m.mul_(0.9).add_(agg)
                        p.add_(m, alpha=-0.1)

                out["ref_ms"] = round(timeit(reference, dev, warmup=2, iters=5), 4)
                out["speedup"] = round(out["ref_ms"] / ours, 2)
                del rows, params, moms, gathered
            if rank == 0:
                print(json.dumps(out), flush=True)
            asyncio.run(ps.shutdown())
            del ps, rnd, nodes
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
