"""Short single-GPU workloads for `ncu` captures of the library's main kernels.

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 -o gpurun_out/<name> \
        python bench/ncu_targets.py <target>

targets: fused_cw (fused PS round kernel, 8 rows, median), gram_tma (tcgen05 Gram fed by TMA, n = 64),
         wsum_multi (one-pass multi-row weighted sum, 64 x 64), bn_cluster (single-launch BatchNorm),
         preagg (n-space NNM map + CAF filter)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from byzpy_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
what = sys.argv[1] if len(sys.argv) > 1 else "fused_cw"
torch.manual_seed(0)
if what == "fused_cw":
    ext = ops.require_ext()
    n, d = 8, 11_689_984
    X = torch.randn(n, d, device=dev)
    agg = torch.zeros(d, device=dev)
    pad = torch.zeros(64, dtype=torch.int32, device=dev)
    ctl = torch.zeros(8, dtype=torch.int32, device=dev)
    params = [torch.randn(d, device=dev) for _ in range(8)]
    moms = [torch.zeros(d, device=dev) for _ in range(8)]
    s = torch.cuda.current_stream().cuda_stream
    for epoch in (1, 2, 3):
        ext.fused_ps_cw([X[i].data_ptr() for i in range(n)], [1.0] * 6 + [-1.0] * 2, ops.MODE_MEDIAN, 0, 0, 0, 0.0, 0.0,
                        d, 0, d, 0, [agg.data_ptr()], [pad.data_ptr()], epoch, 0, ctl.data_ptr(), ctl.data_ptr() + 4,
                        [p.data_ptr() for p in params], [m.data_ptr() for m in moms], 0.05, 0.9, 0.0,
                        ops.sm_count(dev), s)
elif what == "gram_tma":
    X = torch.randn(64, 1 << 24, device=dev)
    for _ in range(3):
        ops.gram(list(X.unbind(0)), impl="umma")
elif what == "wsum_multi":
    X = torch.randn(64, 10_000_000 // 128 * 128, device=dev)
    W = torch.randn(64, 64, device=dev)
    out = torch.empty(64, X.shape[1], device=dev)
    for _ in range(3):
        ops.weighted_sum(list(X.unbind(0)), W, out=out, multi_impl="multi")
elif what == "bn_cluster":
    from byzpy_b200.ops.fused_bn import FusedBatchNorm2d

    x = torch.randn(32, 256, 14, 14, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    bn = FusedBatchNorm2d(256, relu=True).to(dev)
    for _ in range(3):
        y = bn(x)
        y.backward(torch.ones_like(y))
elif what == "preagg":
    from byzpy_b200.ops import nspace_cuda

    G = torch.randn(64, 4096, dtype=torch.float64)
    G = (G @ G.T).to(dev)
    for _ in range(3):
        nspace_cuda.nnm_matrix(G, 16)
        nspace_cuda.caf_coeffs(G, 63, 8, power_iters=3)
torch.cuda.synchronize()
print("done", what)
