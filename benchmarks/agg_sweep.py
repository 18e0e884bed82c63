"""Aggregator throughput sweep (BASELINE.json config 5, single-GPU part).

For n gradient rows of d floats resident on one B200: hand-written kernels vs the reference
operator math executed by PyTorch (`torch.stack` + ATen), device-timed with CUDA events, plus the
achieved fraction of the measured HBM copy peak (MEASURED_PEAKS.json) for the algorithmic bytes.

    python benchmarks/agg_sweep.py --n 8 --dims 1e5,1e6,1e7,1e8 --out gpurun_out/agg_sweep.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from byzpy_b200 import ops  # noqa: E402
from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean  # noqa: E402
from byzpy_b200.aggregators.geometric_wise import GeometricMedian, MultiKrum  # noqa: E402
from byzpy_b200.aggregators.norm_wise import CenteredClipping  # noqa: E402


def peak_gbs() -> float:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, warmup=3, iters=10, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()           # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def ref_median(rows):
    return torch.stack(rows).median(dim=0).values


def ref_trmean(rows, f):
    n = len(rows)
    return torch.stack(rows).sort(dim=0).values[f:n - f].mean(0)


def ref_multikrum(rows, f, q):
    X = torch.stack(rows)
    n = X.shape[0]
    G = X @ X.T
    g = G.diagonal()
    D = (g[:, None] + g[None, :] - 2 * G).clamp_min(0)
    scores = D.sort(dim=1).values[:, 1:n - f].sum(1)
    idx = scores.argsort()[:q]
    return X.index_select(0, idx).mean(0)


def ref_cclip(rows, c_tau, M):
    X = torch.stack(rows)
    v = X.mean(0)
    n = X.shape[0]
    for _ in range(M):
        diff = X - v
        dist = diff.norm(dim=1).clamp_min(1e-12)
        v = v + (diff * torch.minimum(torch.ones_like(dist), c_tau / dist)[:, None]).sum(0) / n
    return v


def ref_gm(rows, tol=1e-6, max_iter=256):
    X = torch.stack(rows)
    z = X.median(0).values
    for _ in range(max_iter):
        dist = (X - z).norm(dim=1).clamp_min(1e-12)
        w = 1 / dist
        z_new = (w[:, None] * X).sum(0) / w.sum()
        if (z_new - z).norm().item() <= tol:
            z = z_new
            break
        z = z_new
    return z


def cw_variants(a):
    """A/B of the three pipelines of csrc/cw_select.cu on the same inputs (bit-identical results are asserted):
    ``python benchmarks/agg_sweep.py --cw-variants --n 64 --dims 1e7`` -- the measurement profiles/cw_select.md
    section 4 is waiting for."""
    from byzpy_b200 import ops

    dev = torch.device("cuda", 0)
    peak = peak_gbs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    results = []
    for ds in a.dims.split(","):
        d = int(float(ds))
        rows = [torch.randn(d, device=dev) for _ in range(a.n)]
        nbytes = (a.n * d + d) * 4
        for mode_name, mode, f in (("median", ops.MODE_MEDIAN, 0), ("trimmed_mean", ops.MODE_TRMEAN, a.f),
                                   ("mean_of_medians", ops.MODE_MEAMED, a.f)):
            base = ops.cw_select(rows, mode, f, impl="direct")
            for impl in ("direct", "staged", "tiled"):
                out = torch.empty(d, device=dev)
                assert torch.equal(ops.cw_select(rows, mode, f, out=out, impl=impl), base), (mode_name, impl)
                t = timeit(lambda: ops.cw_select(rows, mode, f, out=out, impl=impl), flush=flush)
                rec = {"op": mode_name, "impl": impl, "n": a.n, "d": d, "ms": round(t, 4),
                       "frac_of_measured_hbm_peak": round(nbytes / 1e6 / t / peak, 3)}
                results.append(rec)
                print(json.dumps(rec), flush=True)
        del rows
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump({"peak_hbm_gbs": peak, "cw_variants": results}, open(a.out, "w"), indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--f", type=int, default=2)
    ap.add_argument("--dims", default="1e5,1e6,1e7,1e8")
    ap.add_argument("--out", default="gpurun_out/agg_sweep.json")
    ap.add_argument("--skip-ref-above", type=float, default=3e8)
    ap.add_argument("--cw-variants", action="store_true",
                    help="time the coordinate-wise kernel variants (direct / staged / tiled) instead of the operator sweep")
    a = ap.parse_args()
    if a.cw_variants:
        return cw_variants(a)
    dev = torch.device("cuda", 0)
    peak = peak_gbs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    n, f = a.n, a.f
    results = []
    for ds in a.dims.split(","):
        d = int(float(ds))
        rows = [torch.randn(d, device=dev) for _ in range(n)]
        rows[-1].mul_(-1.0)
        bytes_1pass = (n * d + d) * 4
        bytes_2pass = (2 * n * d + d) * 4
        ops_list = [
            ("median", lambda: CoordinateWiseMedian().aggregate(rows), lambda: ref_median(rows), bytes_1pass),
            ("trimmed_mean", lambda: CoordinateWiseTrimmedMean(f=f).aggregate(rows), lambda: ref_trmean(rows, f), bytes_1pass),
            ("multi_krum", lambda: MultiKrum(f=f, q=n - f).aggregate(rows), lambda: ref_multikrum(rows, f, n - f), bytes_2pass),
            ("centered_clipping", lambda: CenteredClipping(c_tau=1.0, M=10).aggregate(rows), lambda: ref_cclip(rows, 1.0, 10), bytes_2pass),
            ("geometric_median", lambda: GeometricMedian(max_iter=64).aggregate(rows), lambda: ref_gm(rows, max_iter=64),
             # pass 1 reads n rows and writes the median start row, pass 2 reads n + 1 rows and writes the result
             bytes_2pass + 3 * d * 4),
        ]
        for name, ours, ref, nbytes in ops_list:
            t_ours = timeit(ours, flush=flush)
            t_ref = None
            if d <= a.skip_ref_above:
                try:
                    t_ref = timeit(ref, warmup=1, iters=3, flush=flush)
                except torch.cuda.OutOfMemoryError:
                    t_ref = None
                    torch.cuda.empty_cache()
            rec = {"op": name, "n": n, "d": d, "ours_ms": round(t_ours, 4),
                   "ref_torch_ms": None if t_ref is None else round(t_ref, 4),
                   "speedup": None if t_ref is None else round(t_ref / t_ours, 2),
                   "algorithmic_GB": round(nbytes / 1e9, 4),
                   "achieved_GBps": round(nbytes / 1e6 / t_ours, 1),
                   "frac_of_measured_hbm_peak": round(nbytes / 1e6 / t_ours / peak, 3)}
            results.append(rec)
            print(json.dumps(rec), flush=True)
        del rows
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump({"peak_hbm_gbs": peak, "results": results}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
