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


# defaults of the reference's per-operator scripts (benchmarks/byzfl/<op>_compare.py): sizes, f
REF_DEFAULTS = {
    "arc": dict(n=256, d=65536, f=8), "bucketing": dict(n=512, d=16384), "caf": dict(n=64, d=65536, f=8),
    "centered-clipping": dict(n=64, d=65536), "clipping": dict(n=256, d=65536),
    "trimmed-mean": dict(n=64, d=65536, f=8), "gaussian": dict(n=64, d=65536), "inf": dict(n=64, d=65536),
    "empire": dict(n=64, d=65536), "little": dict(n=96, d=65536, f=8), "mda": dict(n=18, d=2048, f=6),
    "meamed": dict(n=64, d=65536, f=8), "mimic": dict(n=64, d=65536), "monna": dict(n=64, d=65536, f=8),
    "multi-krum": dict(n=80, d=65536, f=20), "nnm": dict(n=196, d=4096, f=32), "smea": dict(n=12, d=1024, f=3),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", default="median")
    ap.add_argument("--num-grads", "--num-vectors", dest="num_grads", type=int, default=None)
    ap.add_argument("--grad-dim", "--dim", dest="grad_dim", type=int, default=None)
    ap.add_argument("--f", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--timeout", type=float, default=0.0, help="max seconds spent timing ByzFL (0 = no limit)")
    # operator knobs, named as in the reference script of the same operator
    ap.add_argument("--tau", type=float, default=None, help="centered clipping radius / IPM and ALIE scale")
    ap.add_argument("--iters", type=int, default=None)
    ap.add_argument("--threshold", type=float, default=None)
    ap.add_argument("--bucket-size", type=int, default=None)
    ap.add_argument("--mu", type=float, default=None)
    ap.add_argument("--sigma", type=float, default=None)
    ap.add_argument("--epsilon", type=int, default=None)
    ap.add_argument("--reference-index", type=int, default=None)
    ap.add_argument("--q", type=int, default=None)
    a = ap.parse_args()
    ref = REF_DEFAULTS.get(a.op, {})
    a.num_grads = ref.get("n", 64) if a.num_grads is None else a.num_grads
    a.grad_dim = ref.get("d", 65536) if a.grad_dim is None else a.grad_dim
    explicit_f = a.f is not None
    a.f = ref.get("f", 8) if a.f is None else a.f
    # the harness in operator_pool_bench names the same knobs differently for two operators
    a.c_tau = a.tau if a.op == "centered-clipping" else None
    a.scale = -a.tau if (a.op == "empire" and a.tau is not None) else None
    g = torch.Generator().manual_seed(a.seed)
    data = [torch.randn(a.grad_dim, generator=g) for _ in range(a.num_grads)]
    f = a.f if explicit_f else min(a.f, max(0, (a.num_grads - 1) // 2 - 1))
    mk, key = make(a.op, a.num_grads, f, a)
    for _ in range(max(1, a.warmup)):
        direct_call(mk(), key, data)
    t0 = time.perf_counter()
    for _ in range(a.repeat):
        ours = direct_call(mk(), key, data)
    out = {"op": a.op, "n": a.num_grads, "d": a.grad_dim,
           "byzpy_b200_ms": round((time.perf_counter() - t0) / a.repeat * 1e3, 3)}
    byzfl_kwargs = BYZFL[a.op][1](a.num_grads, f) if a.op in BYZFL else {}
    for name, value in (("tau", a.tau), ("c", a.threshold), ("s", a.bucket_size), ("mu", a.mu), ("sigma", a.sigma),
                        ("epsilon", a.epsilon)):
        if value is not None and name in byzfl_kwargs:
            byzfl_kwargs[name] = value
    try:
        import byzfl  # noqa: F401

        cls = getattr(byzfl, BYZFL[a.op][0])
        other = cls(**byzfl_kwargs)
        X = torch.stack(data)
        t0 = time.perf_counter()
        for _ in range(a.repeat):
            theirs = other(X)
            if a.timeout > 0 and time.perf_counter() - t0 > a.timeout:
                break
        out["byzfl_ms"] = round((time.perf_counter() - t0) / a.repeat * 1e3, 3)
        if isinstance(ours, torch.Tensor) and isinstance(theirs, torch.Tensor) and ours.shape == theirs.shape:
            out["max_abs_diff"] = float((ours - theirs).abs().max())
    except Exception as exc:  # noqa: BLE001  (ImportError in the offline image)
        out["byzfl"] = f"unavailable ({type(exc).__name__})"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
