"""Same synthetic inputs through this library and through the external ByzFL library (the reference
ships one ``*_compare.py`` per operator under benchmarks/byzfl/; ByzFL is not installable in the
offline build image, so this single front-end reports ``byzfl: unavailable`` there and still prints
this library's time).

    python benchmarks/byzfl/compare.py --op median --num-grads 64 --grad-dim 65536 [--timeout 120]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from benchmarks.operator_pool_bench import direct_call, make  # noqa: E402

BYZFL = {  # op -> (byzfl class name, ctor kwargs builder)
    "median": ("Median", lambda n, f: {}),
    "trimmed-mean": ("TrMean", lambda n, f: {"f": f}),
    "meamed": ("Meamed", lambda n, f: {"f": f}),
    "multi-krum": ("MultiKrum", lambda n, f: {"f": f}),
    "krum": ("Krum", lambda n, f: {"f": f}),
    "geometric-median": ("GeometricMedian", lambda n, f: {}),
    "mda": ("MDA", lambda n, f: {"f": f}),
    "monna": ("MoNNA", lambda n, f: {"f": f}),
    "smea": ("SMEA", lambda n, f: {"f": f}),
    "centered-clipping": ("CenteredClipping", lambda n, f: {}),
    "caf": ("CAF", lambda n, f: {"f": f}),
    "clipping": ("Clipping", lambda n, f: {"c": 2.0}),
    "arc": ("ARC", lambda n, f: {"f": f}),
    "nnm": ("NNM", lambda n, f: {"f": f}),
    "bucketing": ("Bucketing", lambda n, f: {"s": max(1, n // 16)}),
    # attacks (ByzFL: callables over the stacked honest matrix)
    "gaussian": ("Gaussian", lambda n, f: {"mu": 0.0, "sigma": 1.0}),
    "inf": ("Inf", lambda n, f: {}),
    "empire": ("InnerProductManipulation", lambda n, f: {"tau": 1.0}),   # IPM == Empire with scale -tau
    "little": ("ALittleIsEnough", lambda n, f: {"tau": 1.5}),
    "mimic": ("Mimic", lambda n, f: {"epsilon": 0}),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", default="median")
    ap.add_argument("--num-grads", type=int, default=64)
    ap.add_argument("--grad-dim", type=int, default=65536)
    ap.add_argument("--f", type=int, default=8)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--timeout", type=float, default=120.0)
    a = ap.parse_args()
    g = torch.Generator().manual_seed(0)
    data = [torch.randn(a.grad_dim, generator=g) for _ in range(a.num_grads)]
    f = min(a.f, max(0, (a.num_grads - 1) // 2 - 1))
    mk, key = make(a.op, a.num_grads, f)
    direct_call(mk(), key, data)
    t0 = time.perf_counter()
    for _ in range(a.repeat):
        ours = direct_call(mk(), key, data)
    out = {"op": a.op, "n": a.num_grads, "d": a.grad_dim,
           "byzpy_b200_ms": round((time.perf_counter() - t0) / a.repeat * 1e3, 3)}
    try:
        import byzfl  # noqa: F401

        cls = getattr(byzfl, BYZFL[a.op][0])
        other = cls(**BYZFL[a.op][1](a.num_grads, f))
        X = torch.stack(data)
        t0 = time.perf_counter()
        for _ in range(a.repeat):
            theirs = other(X)
            if time.perf_counter() - t0 > a.timeout:
                break
        out["byzfl_ms"] = round((time.perf_counter() - t0) / a.repeat * 1e3, 3)
        if isinstance(ours, torch.Tensor) and isinstance(theirs, torch.Tensor) and ours.shape == theirs.shape:
            out["max_abs_diff"] = float((ours - theirs).abs().max())
    except Exception as exc:  # noqa: BLE001  (ImportError in the offline image)
        out["byzfl"] = f"unavailable ({type(exc).__name__})"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
