"""Stress harness for the intermittent CPU/GPU disagreement of the synthesised (Little / Empire)
rows: repeats the failing configurations many times in one process, with allocator churn between
repeats, and classifies every disagreement against an fp64 oracle.

    python bench/debug_virtual.py --repeats 2000
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from byzpy_b200 import ops  # noqa: E402
from byzpy_b200.ops import reference as ref  # noqa: E402


def oracle(Xd, virt, f):
    nv, nh, a, b = virt
    H = Xd[:nh]
    v = a * H.mean(0) + b * H.std(0, unbiased=False)
    S = torch.cat([Xd, v[None].expand(nv, -1)]).sort(dim=0).values
    return S[f:S.shape[0] - f].mean(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--repeats", type=int, default=1000)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    if a.threads:
        torch.set_num_threads(a.threads)
    dev = torch.device("cuda", 0)
    virt = (2, 6, 1.0, 1.5)
    bad_runs = 0
    cpu_moves = gpu_moves = 0
    print(f"threads={torch.get_num_threads()} capability={torch.backends.cpu.get_cpu_capability()}", flush=True)
    for it in range(a.repeats):
        d = 5000 if it % 2 == 0 else 4097
        g = torch.Generator().manual_seed(7 + it % 5)
        X = torch.randn(6, d, generator=g)
        rows = [X[i].to(dev).contiguous() for i in range(6)]
        junk = [torch.empty(1 << (10 + (it + k) % 12), device=dev) for k in range(3)]      # allocator churn
        out = ops.cw_select(rows, ops.MODE_TRMEAN, 1, virtual=virt).cpu()
        exp = ref.cw_select([X[i] for i in range(6)], ops.MODE_TRMEAN, 1, virtual=virt)
        col_g = ops.colstat(rows, 1.0, 1.5).cpu()
        col_c = ref.colstat([X[i] for i in range(6)], 1.0, 1.5)
        del junk
        o = oracle(X.double(), virt, 1).float()
        Xd = X.double()
        oc = (Xd.mean(0) + 1.5 * Xd.std(0, unbiased=False)).float()
        for tag, got, want, orc in (("trmean+virtual", out, exp, o), ("colstat", col_g, col_c, oc)):
            if not torch.allclose(got, want, rtol=1e-5, atol=1e-5):
                bad_runs += 1
                bad = ((got - want).abs() > 1e-5 + 1e-5 * want.abs()).nonzero().flatten()
                eg, ec = float((got - orc).abs().max()), float((want - orc).abs().max())
                gpu_moves += eg > 1e-5
                cpu_moves += ec > 1e-5
                print(f"iter {it} {tag} d={d}: {bad.numel()} bad, mod4 {torch.bincount(bad % 4, minlength=4).tolist()}, "
                      f"mod16 {torch.bincount(bad % 16, minlength=16).tolist()}, gpu-vs-fp64 {eg:.2e}, cpu-vs-fp64 {ec:.2e}",
                      flush=True)
    print(f"done: {bad_runs} disagreements in {a.repeats} repeats (gpu side moved {gpu_moves}, cpu side moved {cpu_moves})")


if __name__ == "__main__":
    main()
