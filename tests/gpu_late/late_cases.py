"""GPU tests of kernels written AFTER the round's GPU budget was spent (they have been compiled, statically
analysed and -- where the code is host-compilable -- checked on the CPU, but no B200 has run them yet).  They run in
a CHILD interpreter started by ``tests/test_gpu_zz_round2_late.py`` (which sorts last in the GPU tier): neither a
corrupted CUDA context nor a kernel that never returns can take the results of the measured kernels with it.

Everything exercised here is OPT-IN (``impl="tiled"`` / ``BYZPY_CW_IMPL=tiled``, a pre-aggregator in front of a
coordinate-wise aggregator on the fused path, the example's non-default flags); no default path depends on it.  For
that reason the tests are marked ``xfail(strict=False)``: the first B200 run reports them as XPASS (verified) or XFAIL
(the opt-in feature needs work) without turning the tier of the measured kernels red.  Remove the mark once a run has
shown XPASS."""
import pytest
import torch

from byzpy_b200 import ops

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300),     # (device-side waits have their own 20 s budgets)
              pytest.mark.xfail(strict=False, reason="opt-in code written after the GPU budget was spent: never "
                                                     "executed on a B200 (see the module docstring)")]
DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None


def _rows(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(d, generator=g).to(DEV) for _ in range(n)]


@pytest.mark.parametrize("n", [17, 32, 33, 48, 64, 65, 100, 128])
@pytest.mark.parametrize("mode,f", [(ops.MODE_MEDIAN, 0), (ops.MODE_TRMEAN, 5), (ops.MODE_MEAMED, 6)])
def test_warp_tiled_selection_kernel_is_bit_identical_to_the_direct_kernel(n, mode, f):
    d = 148 * 256 * 4 + 37            # several tiles per warp, plus a tail that is not a whole tile
    rows = _rows(n, d, seed=n)
    want = ops.cw_select(rows, mode, f, impl="direct")
    got = ops.cw_select(rows, mode, f, impl="tiled")
    assert torch.equal(got, want)
    X = torch.stack(rows).cpu()
    if mode == ops.MODE_MEDIAN:
        assert torch.equal(got.cpu(), X.median(dim=0).values)


def test_warp_tiled_kernel_scales_nan_inf_virtual_rows_and_fused_update():
    n, d = 40, 64 * 1024 + 5
    rows = _rows(n, d, seed=7)
    rows[3][::7] = float("nan")
    rows[5][::11] = float("inf")
    rows[6][::13] = float("-inf")
    scales = [1.0] * n
    scales[1], scales[2] = -1.0, 0.5
    for mode, f in [(ops.MODE_MEDIAN, 0), (ops.MODE_TRMEAN, 3)]:
        a = ops.cw_select(rows, mode, f, scales=scales, impl="direct")
        b = ops.cw_select(rows, mode, f, scales=scales, impl="tiled")
        assert torch.equal(torch.nan_to_num(a, nan=123.0), torch.nan_to_num(b, nan=123.0))
    clean = _rows(n, d, seed=8)
    virt = (3, n - 2, 1.0, -0.7)       # three Little-style rows from the first n - 2
    a = ops.cw_select(clean, ops.MODE_MEDIAN, 0, virtual=virt, impl="direct")
    b = ops.cw_select(clean, ops.MODE_MEDIAN, 0, virtual=virt, impl="tiled")
    assert torch.equal(a, b)
    # fused SGD epilogue
    p1, p2 = torch.randn(d, device=DEV), None
    p2 = p1.clone()
    m1, m2 = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    ops.cw_select(clean, ops.MODE_MEDIAN, 0, update=dict(params=[p1], moms=[m1], lr=0.1, momentum=0.9, weight_decay=0.0),
                  impl="direct")
    ops.cw_select(clean, ops.MODE_MEDIAN, 0, update=dict(params=[p2], moms=[m2], lr=0.1, momentum=0.9, weight_decay=0.0),
                  impl="tiled")
    assert torch.equal(p1, p2) and torch.equal(m1, m2)


def test_warp_tiled_kernel_under_graph_capture_and_env_routing(monkeypatch):
    rows = _rows(64, 1 << 20, seed=3)
    out = torch.empty(1 << 20, device=DEV)
    want = ops.cw_select(rows, ops.MODE_MEDIAN, 0, impl="direct")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.cw_select(rows, ops.MODE_MEDIAN, 0, out=out, impl="tiled")
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ops.cw_select(rows, ops.MODE_MEDIAN, 0, out=out, impl="tiled")
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    monkeypatch.setenv("BYZPY_CW_IMPL", "tiled")
    assert torch.equal(ops.cw_select(rows, ops.MODE_MEDIAN, 0), want)


# ------------------------------------------------------------- pre-aggregator -> coordinate-wise fused round
class _TinyNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(20, 33)
        self.b = torch.nn.Linear(33, 5)

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


@pytest.mark.parametrize("pre_name", ["bucketing", "nnm", "clipping", "arc"])
@pytest.mark.parametrize("agg_name", ["median", "trmean"])
@pytest.mark.parametrize("graph", [False, True])
def test_pre_aggregator_with_coordinate_wise_aggregator_runs_fused(pre_name, agg_name, graph):
    """MapCwPlan: Y = W_p X on the coordinate shard, then the fused select / deliver / SGD kernel over the m mixed
    rows -- compared with the host operators applied to autograd gradients of mirror models."""
    import asyncio

    from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean
    from byzpy_b200.engine.node.device import DeviceHonestNode
    from byzpy_b200.engine.parameter_server.ps import ParameterServer
    from byzpy_b200.parallel.device_ps import MapCwPlan
    from byzpy_b200.pre_aggregators import ARC, Bucketing, Clipping, NearestNeighborMixing

    pres = {"clipping": lambda: Clipping(0.05), "arc": lambda: ARC(2), "nnm": lambda: NearestNeighborMixing(2),
            "bucketing": lambda: Bucketing(2, perm=[5, 0, 3, 1, 7, 2, 6, 4])}
    aggs = {"median": lambda: CoordinateWiseMedian(), "trmean": lambda: CoordinateWiseTrimmedMean(f=1)}
    torch.manual_seed(2)
    init = _TinyNet().state_dict()

    def mk():
        net = _TinyNet()
        net.load_state_dict(init)
        return net

    hon = [DeviceHonestNode(mk(), lr=0.1, momentum=0.9, device=DEV) for _ in range(8)]
    ps = ParameterServer(hon, [], aggs[agg_name](), pre_aggregator=pres[pre_name](), fused=True, amp_dtype=None,
                         use_cuda_graph=graph)
    rnd = ps.device_round
    assert isinstance(rnd.plan, MapCwPlan) and rnd.plan.capturable and rnd.use_cuda_graph == graph
    models = [mk().to(DEV) for _ in range(8)]
    lossf = torch.nn.CrossEntropyLoss()
    opts = None
    for t in range(3):
        batches = [(torch.randn(16, 20).pin_memory(), torch.randint(0, 5, (16,)).pin_memory()) for _ in range(8)]
        ps.step(batches)
        rows = []
        for mdl, (x, y) in zip(models, batches):
            mdl.zero_grad()
            lossf(mdl(x.to(DEV)), y.to(DEV)).backward()
            rows.append(torch.cat([p.grad.reshape(-1) for p in mdl.parameters()]).cpu())
        expect = aggs[agg_name]().aggregate(list(pres[pre_name]().pre_aggregate(rows)))
        rnd.read_losses()
        torch.testing.assert_close(rnd.aggregated().cpu(), expect, rtol=2e-4, atol=2e-5)
        for mdl in models:
            off = 0
            for p in mdl.parameters():
                p.grad.copy_(expect[off:off + p.numel()].view_as(p).to(DEV))
                off += p.numel()
        if opts is None:
            opts = [torch.optim.SGD(mdl.parameters(), lr=0.1, momentum=0.9) for mdl in models]
        for o in opts:
            o.step()
    got = hon[0].worker.arena.flat_params[: rnd.d].cpu()
    want = torch.cat([p.detach().reshape(-1) for p in models[0].parameters()]).cpu()
    torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-4)
    asyncio.run(ps.shutdown())


@pytest.mark.parametrize("extra", [[], ["--aggregator", "multikrum", "--pre", "nnm"],
                                   ["--aggregator", "trmean", "--pre", "bucketing", "--timeline", "--buckets", "1"]])
def test_device_example_runs_on_one_gpu(extra):
    """examples/ps/device/resnet_fused.py end to end (64 x 64 images -- the shape smoke() uses -- 6 rounds)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    cmd = [sys.executable, "examples/ps/device/resnet_fused.py", "--rounds", "6", "--image", "64", "--classes", "10",
           "--batch", "4"] + extra
    res = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=280,
                         env=dict(os.environ, PYTHONPATH=root, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0")))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    assert "round 5: losses" in res.stdout and "plan " in res.stdout, res.stdout[-1500:]
    if "--timeline" in extra:
        assert '"round_done"' in res.stdout
