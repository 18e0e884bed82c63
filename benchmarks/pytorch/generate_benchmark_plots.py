"""Bar plots (speed-up over the direct call) of the ActorPool benchmarks.

    python benchmarks/pytorch/generate_benchmark_plots.py --output-dir benchmarks/plots     # runs the benchmarks
    python benchmarks/pytorch/profile_summary.py --out bench_summary.json
    python benchmarks/pytorch/generate_benchmark_plots.py bench_summary.json --out bench_summary.png

Without a summary file the script runs the benchmarks the reference's plots show (MDA, trimmed mean, mean of medians:
reference benchmarks/pytorch/generate_benchmark_plots.py) through profile_summary.py and writes
``<output-dir>/actor_pool_speedups.png`` plus the JSON it was drawn from.
"""
import argparse
import json
import os
import subprocess
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("summary", nargs="?", default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--output-dir", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "plots"))
    ap.add_argument("--ops", default="mda,cwtm,meamed")
    ap.add_argument("--workers", default="2,4,6")
    ap.add_argument("--num-grads", type=int, default=None)
    ap.add_argument("--grad-dim", type=int, default=None)
    ap.add_argument("--pool-backend", default="process")
    a = ap.parse_args()
    if a.summary is None:
        os.makedirs(a.output_dir, exist_ok=True)
        a.summary = os.path.join(a.output_dir, "actor_pool_speedups.json")
        cmd = [sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "profile_summary.py"),
               "--ops", a.ops, "--workers", a.workers, "--pool-backend", a.pool_backend, "--out", a.summary]
        if a.num_grads is not None:
            cmd += ["--num-grads", str(a.num_grads)]
        if a.grad_dim is not None:
            cmd += ["--grad-dim", str(a.grad_dim)]
        subprocess.run(cmd, check=True)
        a.out = a.out or os.path.join(a.output_dir, "actor_pool_speedups.png")
    a.out = a.out or "bench_summary.png"
    rows = [r for r in json.load(open(a.summary)) if "direct_ms" in r]
    try:
        import matplotlib

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        print("matplotlib is not installed; speed-ups as text instead:")
        for r in rows:
            print(r["op"], {k: round(r["direct_ms"] / v, 2) for k, v in r.items() if k.startswith("pool_x")})
        return
    pools = sorted({k for r in rows for k in r if k.startswith("pool_x")})
    fig, ax = plt.subplots(figsize=(max(6, len(rows)), 4))
    w = 0.8 / max(1, len(pools))
    for j, k in enumerate(pools):
        ax.bar([i + j * w for i in range(len(rows))], [r["direct_ms"] / r.get(k, float("nan")) for r in rows],
               width=w, label=k.replace("_ms", ""))
    ax.set_xticks([i + 0.4 - w / 2 for i in range(len(rows))])
    ax.set_xticklabels([r["op"] for r in rows], rotation=45, ha="right")
    ax.set_ylabel("speed-up over direct call")
    ax.axhline(1.0, color="k", lw=0.5)
    ax.legend()
    fig.tight_layout()
    fig.savefig(a.out, dpi=150)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
