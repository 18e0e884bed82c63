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

Single-GPU caveat found with this harness (``--sleep-after-first-rank``): with CUDA's lazy module
loading, the FIRST launch of a kernel the process has not used yet needs a context-wide
synchronisation, so launching such a kernel on the late rank's stream while the early rank's kernels
are already resident and spinning on the late rank's flags never completes (the early rank times out
after its spin budget).  It only exists when both ends of a flag wait live on ONE device; with one
rank per GPU the peer's progress never depends on this process's host thread, and captured graphs
load every kernel at capture time.  The harness therefore warms the spin kernel up first.
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
    ap.add_argument("--no-sleep", action="store_true")
    ap.add_argument("--fixed-order", action="store_true")
    ap.add_argument("--interleave", action="store_true",
                    help="enqueue the ranks' bucket launches alternately instead of rank by rank")
    ap.add_argument("--spin", type=float, default=3.0, help="flag-wait budget in seconds")
    ap.add_argument("--sleep-after-first-rank", action="store_true",
                    help="enqueue the late rank's spin kernel AFTER the first rank's launches (host order)")
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
    if not a.sleep_after_first_rank:
        torch.cuda._sleep(1000)             # load the spin kernel while nothing is waiting on this device
        torch.cuda.synchronize()
    for epoch in range(1, a.epochs + 1):
        X = torch.randn(n, d, device=dev)
        torch.cuda.synchronize()
        order = [0, 1]
        if not a.fixed_order:
            rng.shuffle(order)
        slept = False
        cycles = 0
        if not a.no_sleep and rng.random() < 0.7:
            cycles = int(rng.uniform(2e5, 3e6))
            slept = True
        if cycles and not a.sleep_after_first_rank:
            with torch.cuda.stream(streams[order[1]]):
                torch.cuda._sleep(cycles)                             # the late rank

        def launch(r, k):
            off, ln = bounds[k + 1], bounds[k] - bounds[k + 1]
            half = (ln // 2) // 4 * 4
            s_off = off + r * half
            s_len = half if r == 0 else ln - half
            ext.fused_ps_cw([X[i].data_ptr() for i in range(n)], [1.0] * n, mode, f, 0, 0, 0.0, 0.0, d,
                            s_off, s_len, r, [t.data_ptr() for t in aggs], [p.data_ptr() for p in pads],
                            epoch, 0, ctls[r].data_ptr(), ctls[r].data_ptr() + 4, [params[r].data_ptr()], [],
                            0.1, 0.0, 0.0, sms, streams[r].cuda_stream, limit, off, ln, nb, k, 0, 0, a.spin)

        if a.interleave:
            for k in range(nb):
                for r in order:
                    launch(r, k)
        else:
            for r in order:
                if cycles and a.sleep_after_first_rank and r == order[1]:
                    with torch.cuda.stream(streams[r]):
                        torch.cuda._sleep(cycles)
                for k in range(nb):
                    launch(r, k)
        torch.cuda.synchronize()
        if any(int(c[1].item()) != 0 for c in ctls):
            print(f"epoch {epoch}: kernel status {[int(c[1].item()) for c in ctls]} order {order} slept {slept} "
                  f"ready flags {[p[:2].tolist() for p in pads]} done flags {[p[16:18].tolist() for p in pads]} "
                  f"counters {[int(c[0].item()) for c in ctls]} (bucket b of epoch e publishes e * {nb} + b)")
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
