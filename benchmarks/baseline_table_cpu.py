"""BASELINE.md section 2 workloads, direct call on CPU, ours vs the unmodified reference (same machine).

    python benchmarks/baseline_table_cpu.py            # prints a markdown table
"""
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

ROWS = [  # (label, module path under aggregators/pre_aggregators/attacks, class, ctor kwargs, n, d, call)
    ("MDA n=30 d=2048 f=10", "aggregators.geometric_wise", "MinimumDiameterAveraging", dict(f=10), 30, 2048, "aggregate"),
    ("SMEA n=16 d=4096 f=5", "aggregators.geometric_wise", "SMEA", dict(f=5), 16, 4096, "aggregate"),
    ("ARC n=256 d=65536 f=8", "pre_aggregators", "ARC", dict(f=8), 256, 65536, "pre_aggregate"),
    ("TrimmedMean n=64 d=65536 f=8", "aggregators.coordinate_wise", "CoordinateWiseTrimmedMean", dict(f=8), 64, 65536, "aggregate"),
    ("Median n=64 d=65536", "aggregators.coordinate_wise", "CoordinateWiseMedian", dict(), 64, 65536, "aggregate"),
    ("MultiKrum n=80 d=65536 f=20 q=12", "aggregators.geometric_wise", "MultiKrum", dict(f=20, q=12), 80, 65536, "aggregate"),
    ("GeometricMedian n=64 d=65536", "aggregators.geometric_wise", "GeometricMedian", dict(), 64, 65536, "aggregate"),
    ("CAF n=64 d=65536 f=8", "aggregators.norm_wise", "CAF", dict(f=8), 64, 65536, "aggregate"),
    ("MoNNA n=64 d=65536 f=8", "aggregators.geometric_wise", "MoNNA", dict(f=8), 64, 65536, "aggregate"),
    ("CenteredClipping n=64 d=65536", "aggregators.norm_wise", "CenteredClipping", dict(c_tau=0.1, M=10), 64, 65536, "aggregate"),
    ("CGE n=64 d=65536 f=8", "aggregators.norm_wise", "ComparativeGradientElimination", dict(f=8), 64, 65536, "aggregate"),
    ("MeanOfMedians n=64 d=65536 f=8", "aggregators.coordinate_wise", "MeanOfMedians", dict(f=8), 64, 65536, "aggregate"),
    ("NNM n=196 d=4096 f=32", "pre_aggregators", "NearestNeighborMixing", dict(f=32), 196, 4096, "pre_aggregate"),
    ("Bucketing n=512 d=16384 s=32", "pre_aggregators", "Bucketing", dict(bucket_size=32), 512, 16384, "pre_aggregate"),
    ("Clipping n=256 d=65536 tau=2", "pre_aggregators", "Clipping", dict(threshold=2.0), 256, 65536, "pre_aggregate"),
]


def run(pkg, mod, cls, kw, data, call, budget_s=1.0, min_repeat=5):
    """Median of per-call times: at least ``min_repeat`` calls, more until ``budget_s`` is spent (the short
    workloads are a few ms, where a mean of three calls on a shared box is mostly noise)."""
    C = getattr(importlib.import_module(f"{pkg}.{mod}"), cls)
    getattr(C(**kw), call)(data)                      # warm-up
    times, start = [], time.perf_counter()
    while len(times) < min_repeat or (time.perf_counter() - start < budget_s and len(times) < 200):
        t0 = time.perf_counter()
        getattr(C(**kw), call)(data)
        times.append((time.perf_counter() - t0) * 1e3)
    times.sort()
    return times[len(times) // 2]


def main():
    out = []
    print("| workload | reference direct ms | ours direct ms | speed-up |\n|---|---|---|---|")
    for label, mod, cls, kw, n, d, call in ROWS:
        g = torch.Generator().manual_seed(0)
        data = [torch.randn(d, generator=g) for _ in range(n)]
        ours = run("byzpy_b200", mod, cls, kw, data, call)
        try:
            ref = run("byzpy", mod, cls, kw, data, call)
        except Exception as exc:  # noqa: BLE001
            ref = float("nan")
            print(f"<!-- reference failed on {label}: {exc!r} -->")
        out.append(dict(workload=label, ref_ms=round(ref, 2), ours_ms=round(ours, 2)))
        print(f"| {label} | {ref:.2f} | {ours:.2f} | {ref / ours:.1f}x |", flush=True)
    with open(os.path.join(ROOT, "profiles", "baseline_table_cpu.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
