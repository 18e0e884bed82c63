"""Bar plots (speed-up over the direct call) from a profile_summary.py JSON dump.

    python benchmarks/pytorch/profile_summary.py --out bench_summary.json
    python benchmarks/pytorch/generate_benchmark_plots.py bench_summary.json --out bench_summary.png
"""
import argparse
import json


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("summary")
    ap.add_argument("--out", default="bench_summary.png")
    a = ap.parse_args()
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
