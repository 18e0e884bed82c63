"""Run a set of benchmark scripts and tabulate their JSON result lines (the reference's
benchmarks/pytorch/profile_summary.py scrapes text with regexes; every script here prints one JSON
line, so the summary is exact).

    python benchmarks/pytorch/profile_summary.py --ops median,trimmed_mean,multi_krum --num-grads 32 --grad-dim 32768
    python benchmarks/pytorch/profile_summary.py --out bench_summary.json
    python benchmarks/pytorch/profile_summary.py --bench krum --bench median --workers 2,4
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def discover():
    return sorted(f[:-len("_actor_pool.py")] for f in os.listdir(HERE)
                  if f.endswith("_actor_pool.py") and f not in ("parameter_server_actor_pool.py",
                                                                "mnist_training_actor_pool.py"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="")
    ap.add_argument("--bench", action="append", default=None,
                    help="run only the benchmarks whose name contains this (repeatable; the reference's flag)")
    ap.add_argument("--num-grads", type=int, default=None, help="default: each script's own (= the reference's)")
    ap.add_argument("--grad-dim", type=int, default=None)
    ap.add_argument("--pool-workers", "--workers", dest="pool_workers", default="2,4,6")
    ap.add_argument("--pool-backend", default="thread")
    ap.add_argument("--timeout", type=float, default=600)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    ops = [o for o in a.ops.split(",") if o] or discover()
    if a.bench:
        ops = [o for o in ops if any(b in o or b in f"{o}_actor_pool" for b in a.bench)]
    rows = []
    for op in ops:
        script = os.path.join(HERE, f"{op}_actor_pool.py")
        if not os.path.exists(script):
            script = os.path.join(HERE, f"{op}_preagg.py")
        cmd = [sys.executable, script, "--pool-workers", a.pool_workers, "--pool-backend", a.pool_backend]
        if a.num_grads is not None:
            cmd += ["--num-grads", str(a.num_grads)]
        if a.grad_dim is not None:
            cmd += ["--grad-dim", str(a.grad_dim)]
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=a.timeout)
            line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
            rows.append(json.loads(line))
        except Exception as exc:  # noqa: BLE001
            rows.append({"op": op, "error": repr(exc)})
    keys = ["op", "direct_ms", "scheduler_no_pool_ms"] + sorted({k for r in rows for k in r if k.startswith("pool_x")})
    print(" | ".join(keys))
    for r in rows:
        print(" | ".join(str(r.get(k, "-")) for k in keys))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rows, f, indent=2)


if __name__ == "__main__":
    main()
