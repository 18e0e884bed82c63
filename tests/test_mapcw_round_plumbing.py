"""The pre-aggregator -> coordinate-wise fused round (``DeviceRound._launch_mapcw_round``) was written without a
GPU.  Its kernels are the measured ones; what is new is the plumbing -- which pointers, offsets and lengths each
launch gets (shard-relative output rows, row blocks of W, the Gram pass feeding the map kernel).  This test runs
the REAL method on CPU memory against a stand-in for the extension whose functions implement each kernel's contract
in NumPy over raw addresses, and compares the delivered aggregate and the SGD update with the host operators."""
import ctypes
import types

import numpy as np
import pytest
import torch

from byzpy_b200 import ops
from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean
from byzpy_b200.engine.parameter_server.ps import ParameterServer
from byzpy_b200.parallel.device_ps import DeviceRound, MapCwPlan, RowFold, RowLayout
from byzpy_b200.pre_aggregators import ARC, Bucketing, Clipping, NearestNeighborMixing


def view(ptr, count, dtype=np.float32):
    ctype = {np.float32: ctypes.c_float, np.float64: ctypes.c_double, np.int32: ctypes.c_int32}[dtype]
    return np.ctypeslib.as_array((ctype * count).from_address(ptr))


class FakeExt:
    """Contracts of the kernels ``_launch_mapcw_round`` launches (csrc/api.h, fused_ps.h), on host memory."""

    PAD_READY = 0
    WSUM_MULTI_TILE = 128

    def __init__(self, d_pad):
        self.d_pad = d_pad
        self.calls = []

    def _rows(self, rows, scales, off, ln):
        sc = scales if len(scales) else [1.0] * len(rows)
        return np.stack([view(p + 4 * off, ln).astype(np.float64) * s for p, s in zip(rows, sc)])

    def gram_partials_needed(self, n, sm):
        return n * n * 4

    def gram_umma_tile_cols(self, n):
        return 96          # does not divide every padded length: the CUDA-core tail pass is exercised too

    def flag_barrier(self, pads, rank, slot, epoch_ptr, status, stream, seq_mul=0, seq_add=0, live_mask=0, spin_s=0.0):
        self.calls.append("flag_barrier")

    def colstat(self, rows, scales, a, b, off, ln, out, sm, stream):
        self.calls.append("colstat")
        X = self._rows(rows, scales, off, ln)
        view(out + 4 * off, ln)[:] = a * X.mean(0) + b * X.std(0)

    def gram(self, rows, scales, off, ln, partials, num_partials, G, G64, sm, stream, aux_median=0):
        self.calls.append("gram")
        X = self._rows(rows, scales, off, ln)
        n = len(rows)
        g = X @ X.T
        view(G, n * n)[:] = g.reshape(-1)
        if G64:
            view(G64, n * n, np.float64)[:] = g.reshape(-1)

    def gram_umma(self, rows, scales, off, main, partials, slots, tail_ptr, G, G64, sm, stream):
        self.calls.append("gram_umma")
        assert main % 96 == 0
        X = self._rows(rows, scales, off, main)
        n = len(rows)
        g = X @ X.T
        if tail_ptr:
            g = g + view(tail_ptr, n * n, np.float64).reshape(n, n)
        view(G, n * n)[:] = g.reshape(-1)
        if G64:
            view(G64, n * n, np.float64)[:] = g.reshape(-1)

    def gram_exchange(self, local, slots, pads, rank, n, epoch_ptr, status, out64, out32, stream, live_mask=0, spin_s=0.0,
                      slots_mc=0):
        self.calls.append("gram_exchange")
        view(out64, n * n, np.float64)[:] = view(local, n * n, np.float64)

    def _wsum(self, rows, scales, W, m, off, ln, outs):
        n = len(rows)
        Wm = view(W, m * n).reshape(m, n).astype(np.float64)
        Y = Wm @ self._rows(rows, scales, off, ln)
        for r in range(m):
            view(outs[r] + 4 * off, ln)[:] = Y[r]          # outputs are indexed with the GLOBAL coordinate

    def wsum_multi(self, rows, scales, W, m, off, ln, outs, sm, stream):
        self.calls.append("wsum_multi")
        assert m > 8 and ln % self.WSUM_MULTI_TILE == 0 and all(p % 16 == 0 for p in outs) and off % 4 == 0
        self._wsum(rows, scales, W, m, off, ln, outs)

    def wsum(self, rows, scales, W, m, off, ln, outs, upd_params, upd_moms, lr, mu, wd, sm, stream):
        self.calls.append("wsum")
        assert 1 <= m <= 8 and len(outs) == m and not upd_params
        self._wsum(rows, scales, W, m, off, ln, outs)

    def fused_ps_cw(self, rows, scales, mode, f, nv, nh, va, vb, d, shard_off, shard_len, rank, agg, pads, epoch, epoch_ptr,
                    counter, status, upd_params, upd_moms, lr, mu, wd, sm, stream, grid_limit=0, rng_off=0, rng_len=0,
                    nb=0, k=0, agg_mc=0, live_mask=0, spin_s=0.0, trace=0):
        self.calls.append("fused_ps_cw")
        assert nv == 0 and all(p % 16 == 0 for p in rows) and shard_off % 4 == 0 and shard_len % 4 == 0
        assert rng_off == 0 and rng_len == d == self.d_pad and nb == 1 and k == 0
        X = torch.from_numpy(self._rows(rows, scales, shard_off, shard_len).astype(np.float32))
        res = ops.cw_select(list(X.unbind(0)), mode, f)
        view(agg[rank] + 4 * shard_off, shard_len)[:] = res.numpy()
        g = view(agg[rank], d)
        for i, p in enumerate(upd_params):
            pv = view(p, d)
            gg = g + wd * pv
            if upd_moms:
                mv = view(upd_moms[i], d)
                mv[:] = mu * mv + gg
                gg = mv
            pv[:] = pv - lr * gg


class _Dev:
    device = torch.device("cpu")


def _standin(n_rows, d, plan, n_virtual=0, n_honest=None, fold=None):
    """A DeviceRound without __init__: exactly the state _setup_mapcw_plan / _launch_mapcw_round read."""
    d_pad = (d + 1023) // 1024 * 1024
    n_workers = n_rows - n_virtual
    r = object.__new__(DeviceRound)
    r.ext = FakeExt(d_pad)
    r.device = torch.device("cpu")
    r.world, r.rank, r.live_mask = 1, 0, 0
    r.d, r.d_pad, r.sm, r.nt_max = d, d_pad, 8, 144
    r.layout = RowLayout(n_honest if n_honest is not None else n_workers, n_workers - (n_honest or n_workers), n_virtual, 1,
                         [0] * n_workers, list(range(n_workers)))
    r.plan = plan
    r.virtual_fold = fold
    r.spin_seconds = 0.0
    g = torch.Generator().manual_seed(11)
    r._grads_t = torch.randn(n_workers, d_pad, generator=g)
    r._grads_t[:, d:] = 0.0
    r._agg_t = torch.zeros(d_pad)
    r._pad_t = torch.zeros(64, dtype=torch.int32)
    r._gslots_t = torch.zeros(2 * 144 * 144, dtype=torch.float64)
    r._ctl_t = torch.zeros(64, dtype=torch.int32)
    r._rows = [r._grads_t[i].data_ptr() for i in range(n_workers)]
    r._scales = [1.0] * n_workers
    r._off_pad, r._off_agg, r._off_gslots = 1, 2, 3
    table = {1: r._pad_t.data_ptr(), 2: r._agg_t.data_ptr(), 3: r._gslots_t.data_ptr()}
    r.sym = types.SimpleNamespace(peer_ptr=lambda rank, off: table[off], mc_ptr=lambda off: 0)
    r._agg_mc = 0
    r._params_t = torch.randn(2, d_pad, generator=g)
    r._moms_t = torch.zeros(2, d_pad)
    r._upd_params = [r._params_t[i].data_ptr() for i in range(2)]
    r._upd_moms = [r._moms_t[i].data_ptr() for i in range(2)]
    r.lr, r.momentum, r.weight_decay = 0.1, 0.9, 0.01
    r._setup_mapcw_plan()
    return r


PRES = {"bucketing": lambda: Bucketing(2, perm=[5, 0, 3, 1, 7, 2, 6, 4, 9, 8]), "nnm": lambda: NearestNeighborMixing(2),
        "clipping": lambda: Clipping(20.0), "arc": lambda: ARC(2),
        "bucketing20": lambda: Bucketing(2, perm=list(range(19, -1, -1)))}


@pytest.mark.parametrize("pre_name,n", [("bucketing", 10), ("nnm", 10), ("clipping", 10), ("arc", 10), ("nnm", 20),
                                        ("bucketing20", 20)])
@pytest.mark.parametrize("agg_name", ["median", "trmean"])
@pytest.mark.parametrize("d", [3000, 5000 + 37])
def test_map_round_plumbing_against_host_operators(pre_name, n, agg_name, d):
    mk_agg = {"median": CoordinateWiseMedian, "trmean": lambda: CoordinateWiseTrimmedMean(f=1)}[agg_name]
    ps = ParameterServer([_Dev()], [], mk_agg(), pre_aggregator=PRES[pre_name](), fused=None)
    plan = ps._fused_plan(n)
    assert isinstance(plan, MapCwPlan)
    if plan.refresh is not None:
        plan.refresh()
    r = _standin(n, d, plan)
    params0 = r._params_t.clone()
    r._launch_mapcw_round(0, r._ctl_t.data_ptr())
    rows = [r._grads_t[i, :d].clone() for i in range(n)]
    expect = mk_agg().aggregate(list(PRES[pre_name]().pre_aggregate(rows)))
    torch.testing.assert_close(r._agg_t[:d], expect, rtol=1e-4, atol=1e-5)
    assert float(r._agg_t[d:].abs().max()) == 0.0                       # the zero padding aggregates to zero
    g = r._agg_t + 0.01 * params0[0]
    torch.testing.assert_close(r._params_t[0], params0[0] - 0.1 * g, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(r._moms_t[1], r._agg_t + 0.01 * params0[1], rtol=1e-5, atol=1e-6)
    calls = r.ext.calls
    assert calls[0] == "flag_barrier" and calls[-1] == "fused_ps_cw"
    assert ("gram_exchange" in calls) == plan.needs_gram
    assert ("wsum_multi" in calls) == (plan.m > 8)                      # 20 rows -> the one-pass kernel + a tail


def test_map_round_with_virtual_little_rows():
    n_workers, n_virtual, d = 8, 2, 2500
    ps = ParameterServer([_Dev()], [], CoordinateWiseMedian(), pre_aggregator=Bucketing(2, perm=list(range(10))), fused=None)
    plan = ps._fused_plan(n_workers + n_virtual)
    plan.refresh()
    fold = RowFold("virtual", a=1.0, b=-0.8)
    r = _standin(n_workers + n_virtual, d, plan, n_virtual=n_virtual, n_honest=n_workers, fold=fold)
    r._launch_mapcw_round(0, r._ctl_t.data_ptr())
    rows = [r._grads_t[i, :d].clone() for i in range(n_workers)]
    X = torch.stack(rows).double()
    virt = (1.0 * X.mean(0) - 0.8 * X.std(0, unbiased=False)).float()
    mixed = Bucketing(2, perm=list(range(10))).pre_aggregate(rows + [virt, virt])
    torch.testing.assert_close(r._agg_t[:d], CoordinateWiseMedian().aggregate(list(mixed)), rtol=1e-4, atol=1e-5)
    assert "colstat" in r.ext.calls
