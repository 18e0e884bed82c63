"""Diagnostic for the intermittent LittleAttack CPU/GPU mismatch seen in tests/test_gpu_engine.py:
repeats the comparison many times in the process state the test runs in (after the aggregator and
pre-aggregator comparisons) and, on any mismatch, reports which side deviates from an fp64 oracle,
where the bad elements are, and whether the inputs themselves arrived intact on the device.

    python bench/debug_little.py > gpurun_out/little_debug.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from byzpy_b200 import ops  # noqa: E402
from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian  # noqa: E402
from byzpy_b200.aggregators.geometric_wise import GeometricMedian, MultiKrum  # noqa: E402
from byzpy_b200.attacks import EmpireAttack, LittleAttack  # noqa: E402
from byzpy_b200.pre_aggregators import ARC, Clipping, NearestNeighborMixing  # noqa: E402

DEV = torch.device("cuda", 0)


def grads(n, d, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(d, generator=g) + (4.0 if i < 2 else 0.0) for i in range(n)]


def runs(idx):
    """Compress sorted indices into [start, end] runs."""
    out, start, prev = [], None, None
    for i in idx:
        if start is None:
            start = prev = i
        elif i == prev + 1:
            prev = i
        else:
            out.append((start, prev))
            start = prev = i
    if start is not None:
        out.append((start, prev))
    return out


def main():
    print("torch", torch.__version__, torch.cuda.get_device_name(0), "threads", torch.get_num_threads())
    # the state the failing test runs in: the two test functions that precede it
    g11 = grads(11, 3001, seed=4)
    for mk in (lambda: CoordinateWiseMedian(), lambda: MultiKrum(f=2, q=3), lambda: GeometricMedian()):
        mk().aggregate([x.to(DEV) for x in g11])
    g11 = grads(11, 2049, seed=5)
    for mk in (lambda: Clipping(threshold=30.0), lambda: ARC(f=3), lambda: NearestNeighborMixing(f=3)):
        mk().pre_aggregate([x.to(DEV) for x in g11])

    g = grads(7, 4097, seed=6)
    X64 = torch.stack(g).double()
    z = LittleAttack(f=2).z_value(7)
    oracle = X64.mean(0) + z * X64.std(0, unbiased=False)
    bad_runs = 0
    for it in range(400):
        fresh = it % 2 == 0
        gd = [x.to(DEV) for x in g] if fresh else gd_keep
        if it == 0:
            gd_keep = [x.to(DEV) for x in g]
        cpu = LittleAttack(f=2).apply(honest_grads=g)
        gpu = LittleAttack(f=2).apply(honest_grads=gd).cpu()
        e_cpu = (cpu.double() - oracle).abs()
        e_gpu = (gpu.double() - oracle).abs()
        if e_cpu.max() > 1e-5 or e_gpu.max() > 1e-5:
            bad_runs += 1
            side = "CPU" if e_cpu.max() > e_gpu.max() else "GPU"
            err = e_cpu if side == "CPU" else e_gpu
            idx = torch.nonzero(err > 1e-5).reshape(-1).tolist()
            intact = [bool(torch.equal(r.cpu(), x)) for r, x in zip(gd, g)]
            mean_gpu = EmpireAttack(scale=1.0).apply(honest_grads=gd).cpu()
            e_mean = (mean_gpu.double() - X64.mean(0)).abs().max().item()
            std_t = torch.stack(gd).double().std(0, unbiased=False).cpu()
            e_std_t = (std_t - X64.std(0, unbiased=False)).abs().max().item()
            again = LittleAttack(f=2).apply(honest_grads=gd).cpu()
            print(f"iter {it} fresh={fresh} side={side} max_err={err.max().item():.3e} n_bad={len(idx)} "
                  f"runs={runs(idx)[:12]} inputs_intact={intact} gpu_mean_err={e_mean:.2e} "
                  f"torch_std_on_device_err={e_std_t:.2e} rerun_err={(again.double() - oracle).abs().max().item():.2e}")
            if bad_runs >= 10:
                break
    print("bad runs:", bad_runs, "launch counter", ops.launches())


if __name__ == "__main__":
    main()
