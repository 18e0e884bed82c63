"""Tiny driver for single-kernel ncu captures of the stand-alone operators.

    ncu --set full --clock-control none --import-source on -k regex:cw_select -s 2 -c 1 \
        -o gpurun_out/trmean_n8 python bench/ncu_ops.py --op trmean --n 8 --d 1e8
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from byzpy_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--op", default="trmean", choices=["median", "trmean", "meamed", "gram", "gram_umma", "wsum"])
ap.add_argument("--n", type=int, default=8)
ap.add_argument("--d", type=float, default=1e8)
ap.add_argument("--f", type=int, default=2)
ap.add_argument("--iters", type=int, default=4)
a = ap.parse_args()
d = int(a.d)
torch.manual_seed(0)
rows = [torch.randn(d, device="cuda") for _ in range(a.n)]
out = torch.empty(d, device="cuda")
for _ in range(a.iters):
    if a.op == "median":
        ops.cw_median(rows, out=out)
    elif a.op == "trmean":
        ops.cw_trimmed_mean(rows, a.f, out=out)
    elif a.op == "meamed":
        ops.cw_meamed(rows, a.f, out=out)
    elif a.op == "gram":
        ops.gram(rows, impl="fp32")
    elif a.op == "gram_umma":
        ops.gram(rows, impl="umma")
    elif a.op == "wsum":
        w = torch.full((1, a.n), 1.0 / a.n, device="cuda")
        ops.weighted_sum(rows, w)
torch.cuda.synchronize()
print("ok")
