"""Gram kernel micro-benchmark: tcgen05 (3xTF32) vs exact fp32 CUDA-core vs cuBLAS (torch.mm on the
stacked matrix, TF32 disabled), for n in {8..128}; reports ms, achieved GB/s and TFLOP/s."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from byzpy_b200 import ops  # noqa: E402
from benchmarks.agg_sweep import peak_gbs, timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ns", default="8,16,32,64,128")
    ap.add_argument("--d", type=float, default=2 ** 24)
    ap.add_argument("--out", default="gpurun_out/gram_bench.json")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    d = int(a.d)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak = peak_gbs()
    out = []
    for n in [int(x) for x in a.ns.split(",")]:
        dd = min(d, int(6e9 // (4 * n)))
        dd -= dd % 64
        X = torch.randn(n, dd, device=dev)
        rows = [X[i] for i in range(n)]
        nbytes = n * dd * 4
        flops = 2.0 * n * n * dd
        rec = {"n": n, "d": dd}
        for impl in ("umma", "fp32"):
            t = timeit(lambda: ops.gram(rows, impl=impl), flush=flush)
            rec[f"{impl}_ms"] = round(t, 4)
            rec[f"{impl}_GBps"] = round(nbytes / 1e6 / t, 1)
            rec[f"{impl}_frac_hbm"] = round(nbytes / 1e6 / t / peak, 3)
            rec[f"{impl}_TFLOPs"] = round(flops / 1e9 / t, 2)
        torch.backends.cuda.matmul.allow_tf32 = False
        t = timeit(lambda: X @ X.T, flush=flush)
        rec["cublas_fp32_ms"] = round(t, 4)
        torch.backends.cuda.matmul.allow_tf32 = True
        t = timeit(lambda: X @ X.T, flush=flush)
        rec["cublas_tf32_ms"] = round(t, 4)
        out.append(rec)
        print(json.dumps(rec), flush=True)
        del X, rows
        torch.cuda.empty_cache()
    json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
