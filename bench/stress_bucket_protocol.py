"""Stress test of the bucketed cross-rank flag protocol of the fused parameter-server kernel, with
both "ranks" on ONE GPU (two streams, two signal pads), so it can run under compute-sanitizer:

    python bench/stress_bucket_protocol.py --epochs 200
    compute-sanitizer --tool racecheck  python bench/stress_bucket_protocol.py --epochs 20
    compute-sanitizer --tool synccheck  python bench/stress_bucket_protocol.py --epochs 20
    compute-sanitizer --tool memcheck   python bench/stress_bucket_protocol.py --epochs 20

Every epoch launches, per rank, one kernel per gradient bucket (reverse order, per-bucket sequence
numbers epoch * nb + bucket in the flag words); the ranks' launch order is randomised and one of them
is delayed by a spin kernel, so ready / delivery waits really wait.  After every epoch the
aggregate in both ranks' buffers and the SGD-updated parameters are compared with a torch oracle.
"""
import argparse
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from byzpy_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--d", type=int, default=1 << 16)
    ap.add_argument("--buckets", type=int, default=4)
    ap.add_argument("--mode", default="median", choices=["median", "trmean"])
    a = ap.parse_args()
    ext = ops.require_ext()
    dev = torch.device("cuda", 0)
    n, d, nb, world = 8, a.d, a.buckets, 2
    sms = ops.sm_count(dev)
    rng = random.Random(0)
    aggs = [torch.zeros(d, device=dev) for _ in range(world)]
    pads = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(world)]
    ctls = [torch.zeros(8, dtype=torch.int32, device=dev) for _ in range(world)]
    params = [torch.zeros(d, device=dev) for _ in range(world)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    bounds = [d] + [d - (k + 1) * (d // nb) // 1024 * 1024 for k in range(nb - 1)] + [0]
    mode, f = (ops.MODE_MEDIAN, 0) if a.mode == "median" else (ops.MODE_TRMEAN, 2)
    limit = max(1, sms // 4)
    expect_params = torch.zeros(d, device=dev)
    bad = 0
    for epoch in range(1, a.epochs + 1):
        X = torch.randn(n, d, device=dev)
        torch.cuda.synchronize()
        order = [0, 1]
        rng.shuffle(order)
        for r in order:
            with torch.cuda.stream(streams[r]):
                if r == order[1] and rng.random() < 0.7:
                    torch.cuda._sleep(int(rng.uniform(2e5, 3e6)))     # the late rank
                for k in range(nb):
                    off, ln = bounds[k + 1], bounds[k] - bounds[k + 1]
                    half = (ln // 2) // 4 * 4
                    s_off = off + r * half
                    s_len = half if r == 0 else ln - half
                    ext.fused_ps_cw([X[i].data_ptr() for i in range(n)], [1.0] * n, mode, f, 0, 0, 0.0, 0.0, d,
                                    s_off, s_len, r, [t.data_ptr() for t in aggs], [p.data_ptr() for p in pads],
                                    epoch, 0, ctls[r].data_ptr(), ctls[r].data_ptr() + 4, [params[r].data_ptr()], [],
                                    0.1, 0.0, 0.0, sms, streams[r].cuda_stream, limit, off, ln, nb, k)
        torch.cuda.synchronize()
        if any(int(c[1].item()) != 0 for c in ctls):
            print(f"epoch {epoch}: kernel status {[int(c[1].item()) for c in ctls]}")
            bad += 1
            break
        exp = X.median(dim=0).values if a.mode == "median" else X.sort(dim=0).values[2:6].mean(dim=0)
        expect_params -= 0.1 * exp
        for r in range(world):
            if not torch.allclose(aggs[r], exp, rtol=1e-6, atol=1e-6):
                bad += 1
                print(f"epoch {epoch}: rank {r} aggregate mismatch {(aggs[r] - exp).abs().max().item():.3e}")
            if not torch.allclose(params[r], expect_params, rtol=1e-5, atol=1e-5):
                bad += 1
                print(f"epoch {epoch}: rank {r} parameter mismatch {(params[r] - expect_params).abs().max().item():.3e}")
        want = epoch * nb + nb - 1
        for r in range(world):
            flags = pads[r].cpu()
            assert int(flags[0]) == want and int(flags[1]) == want and int(flags[16]) == want and int(flags[17]) == want, flags[:20]
    print(f"STRESS_BUCKET_PROTOCOL {'PASS' if bad == 0 else 'FAIL'} epochs={a.epochs} buckets={nb} mode={a.mode}")
    sys.exit(0 if bad == 0 else 1)


if __name__ == "__main__":
    main()
