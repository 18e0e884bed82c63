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
    ps._allow_mapcw = True
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
    ps._allow_mapcw = True
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


# --------------------------------------------------------------------------- Gram round (GPU-validated) on host memory
def _fake_gram_extras(ext):
    """The two further kernels the Gram round launches."""

    def cw_select(rows, scales, mode, f, nv, nh, va, vb, off, ln, out, upd_params, upd_moms, lr, mu, wd, sm, stream, impl=0):
        ext.calls.append("cw_select")
        X = torch.from_numpy(ext._rows(rows, scales, off, ln).astype(np.float32))
        view(out + 4 * off, ln)[:] = ops.cw_select(list(X.unbind(0)), mode, f).numpy()

    def fused_ps_wsum(rows, scales, W, d, shard_off, shard_len, rank, agg, pads, epoch_ptr, counter, status, upd_params,
                      upd_moms, lr, mu, wd, sm, stream, grid_limit=0, rng_off=0, rng_len=0, seq_mul=0, seq_add=0,
                      agg_mc=0, live_mask=0, spin_s=0.0):
        ext.calls.append("fused_ps_wsum")
        n = len(rows)
        w = view(W, n).astype(np.float64)
        view(agg[rank] + 4 * shard_off, shard_len)[:] = w @ ext._rows(rows, scales, shard_off, shard_len)
        ext.covered = (shard_off, shard_len)

    ext.cw_select, ext.fused_ps_wsum = cw_select, fused_ps_wsum


@pytest.mark.parametrize("live_mask,world,rank", [(0, 1, 0), (0b01, 2, 0), (0, 2, 1), (0b1101, 4, 2)])
def test_gram_round_shard_follows_the_live_set(live_mask, world, rank):
    """``_launch_gram_round`` on host memory: with every rank alive a rank aggregates its static share; after peers
    were dropped (``recover()``) the survivors' shares tile the whole vector (a single survivor covers all of it)."""
    from byzpy_b200.aggregators.geometric_wise import GeometricMedian, MultiKrum

    n, d = 6, 4000
    for mk in (lambda: MultiKrum(f=1, q=3), lambda: GeometricMedian(tol=1e-7)):
        plan = mk().fused_plan(n)
        r = object.__new__(DeviceRound)
        d_pad = 4096
        r.ext = FakeExt(d_pad)
        _fake_gram_extras(r.ext)
        r.device = torch.device("cpu")
        r.world, r.rank, r.live_mask = world, rank, live_mask
        r.d, r.d_pad, r.sm, r.nt_max = d, d_pad, 8, 144
        r.layout = RowLayout(n, 0, 0, world, [0] * n, list(range(n)))
        r.plan, r.virtual_fold, r.spin_seconds = plan, None, 0.0
        g = torch.Generator().manual_seed(5)
        r._grads_t = torch.randn(n, d_pad, generator=g)
        r._grads_t[:, d:] = 0.0
        r._agg_t = torch.zeros(d_pad)
        r._pad_t = torch.zeros(64, dtype=torch.int32)
        r._gslots_t = torch.zeros(world * 144 * 144 + 144 * 144, dtype=torch.float64)
        r._ctl_t = torch.zeros(64, dtype=torch.int32)
        r._rows = [r._grads_t[i].data_ptr() for i in range(n)]
        r._scales = [1.0] * n
        r._off_pad, r._off_agg, r._off_gslots = 1, 2, 3
        table = {1: r._pad_t.data_ptr(), 2: r._agg_t.data_ptr(), 3: r._gslots_t.data_ptr()}
        r.sym = types.SimpleNamespace(peer_ptr=lambda rk, off: table[off], mc_ptr=lambda off: 0)
        r._agg_mc = 0
        r._upd_params, r._upd_moms = [], []
        r.lr, r.momentum, r.weight_decay = 0.1, 0.0, 0.0
        r._setup_gram_plan()
        r._launch_gram_round(0, r._ctl_t.data_ptr())
        off, ln = r.ext.covered
        live = [k for k in range(world) if live_mask == 0 or (live_mask >> k) & 1]
        share = (d_pad // len(live)) // 4 * 4
        idx = live.index(rank)
        assert off == idx * share and ln == (share if idx < len(live) - 1 else d_pad - share * (len(live) - 1))
        if len(live) == 1:
            # the lone survivor's Gram is the full Gram: its aggregate must be the operator's
            assert (off, ln) == (0, d_pad)
            rows = [r._grads_t[i, :d].clone() for i in range(n)]
            torch.testing.assert_close(r._agg_t[:d], mk().aggregate(rows), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("world,live_mask", [(1, 0), (2, 0), (8, 0), (4, 0b1011), (8, 0b00010001)])
def test_bucket_launches_of_all_ranks_tile_the_arena_exactly_once(world, live_mask):
    """``_launch_cw_bucket`` argument plumbing: over all live ranks and all buckets the (shard_off, shard_len) ranges
    are disjoint, 16-byte aligned and cover the padded arena; every launch of a bucket carries the bucket's range and
    the same sequence numbering."""
    from byzpy_b200.parallel.device_ps import CwPlan, bucket_bounds

    d_pad = 11_689_984
    bounds = bucket_bounds([6_400_000, 2_700_000, 700_000], d_pad, 1 << 16)
    assert bounds[0] == d_pad and bounds[-1] == 0 and len(bounds) == 5
    seen = []

    class Rec:
        PAD_READY = 0

        def fused_ps_cw(self, rows, scales, mode, f, nv, nh, va, vb, d, s_off, s_len, rank, agg, pads, epoch, epoch_ptr,
                        counter, status, up, um, lr, mu, wd, sm, stream, grid_limit, rng_off, rng_len, nb, k, agg_mc,
                        live, spin, trace):
            assert s_off % 4 == 0 and s_len % 4 == 0 and rng_off <= s_off and s_off + s_len <= rng_off + rng_len
            seen.append((rank, k, nb, rng_off, rng_len, s_off, s_len, live))

    live = [r for r in range(world) if live_mask == 0 or (live_mask >> r) & 1]
    for rank in live:
        r = object.__new__(DeviceRound)
        r.ext = Rec()
        r.device = torch.device("cpu")
        r.world, r.rank, r.live_mask = world, rank, live_mask
        r.d_pad, r.sm = d_pad, 148
        r.plan = CwPlan(0, 0)
        r.layout = RowLayout(8, 0, 0, world, [0] * 8, list(range(8)))
        r.virtual_fold = None
        r._bounds = bounds
        r._rows, r._scales = [16 * (i + 1) for i in range(8)], [1.0] * 8
        r._off_agg = r._off_pad = 0
        r.sym = types.SimpleNamespace(peer_ptr=lambda rk, off: 4096 * (rk + 1))
        r.ctl = torch.zeros(8, dtype=torch.int32)
        r._upd_params, r._upd_moms = [], []
        r.lr, r.momentum, r.weight_decay, r._agg_mc, r.spin_seconds, r._trace = 0.1, 0.0, 0.0, 0, 0.0, None
        r._round_launches = 0
        import unittest.mock as mock

        with mock.patch("torch.cuda.current_stream", return_value=types.SimpleNamespace(cuda_stream=0)):
            for k in range(len(bounds) - 1):
                r._launch_cw_bucket(k, 37)
    cover = np.zeros(d_pad // 4, dtype=np.int32)
    for rank, k, nb, rng_off, rng_len, s_off, s_len, lv in seen:
        assert nb == len(bounds) - 1 and (rng_off, rng_len) == (bounds[k + 1], bounds[k] - bounds[k + 1]) and lv == live_mask
        cover[s_off // 4: (s_off + s_len) // 4] += 1
    assert cover.min() == 1 and cover.max() == 1
    assert len(seen) == len(live) * (len(bounds) - 1)
