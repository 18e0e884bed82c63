"""Per-layer timing of the fused BatchNorm kernels on the five ResNet-18 activation shapes (batch 32),
against the bytes each direction must move.  Run under gpurun; prints one line per (shape, direction)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from byzpy_b200.ops.fused_bn import FusedBatchNorm2d  # noqa: E402

dev = torch.device("cuda", 0)
PEAK = 6482.7
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, it=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(it):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3


SHAPES = [(32, 64, 112, 112), (32, 64, 56, 56), (32, 128, 28, 28), (32, 256, 14, 14), (32, 512, 7, 7)]
if "--stem-only" in sys.argv:      # for ncu captures of the big-activation kernels
    SHAPES = SHAPES[:1]
for shape in SHAPES:
    for res in ((False,) if "--stem-only" in sys.argv else (False, True)):
        N, C, H, W = shape
        mk = lambda: torch.randn(shape, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        x, r, gy = mk().requires_grad_(True), mk().requires_grad_(True), mk()
        bn = FusedBatchNorm2d(C, relu=True).to(dev)
        mb = x.numel() * 2 / 1e6
        with torch.no_grad():
            pass
        fwd = timeit(lambda: bn(x, r if res else None))

        def both():          # autograd runs backward on the forward's stream: capture them together
            y = bn(x, r if res else None)
            y.backward(gy)
            x.grad = None
            r.grad = None

        b = timeit(both) - fwd
        fwd_bytes = mb * (3 + (1 if res else 0))            # x twice, y once (+ residual)
        bwd_bytes = mb * (5 + (3 if res else 0))            # x, dy twice, dx (+ y twice, dres)
        print(f"shape={shape} residual={res} act={mb:.1f}MB  fwd {fwd:6.1f} us ({fwd_bytes / fwd / 1e3 / PEAK * 1e3:.2f} of HBM peak)"
              f"  bwd {b:6.1f} us ({bwd_bytes / b / 1e3 / PEAK * 1e3:.2f} of HBM peak)")
