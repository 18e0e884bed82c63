"""Native CPU selection-network kernel (csrc/host_select.cpp) against the plain PyTorch oracle
(ops/reference.py): every mode, odd / even n, ties, NaN / +-inf, per-row scales, thread counts, tile
tails, and the dispatch rules of ``ops.cw_select`` on CPU tensors."""
from __future__ import annotations

import pytest
import torch

from byzpy_b200 import ops
from byzpy_b200.ops import reference as ref

pytestmark = pytest.mark.skipif(not ops.extension_available(), reason="kernel library not built")

MODES = [(ops.MODE_MEDIAN, "median"), (ops.MODE_TRMEAN, "trmean"), (ops.MODE_MEAMED, "meamed"),
         (ops.MODE_MEAN, "mean")]


def _host(rows, mode, f=0, scales=(), threads=4):
    from byzpy_b200 import _C

    out = torch.full((rows[0].numel(),), float("nan"))
    _C.host_cw_select([r.data_ptr() for r in rows], list(scales), mode, f, rows[0].numel(), out.data_ptr(), threads)
    return out


def _same(a, b, tol=2e-6):
    fa = torch.nan_to_num(a, nan=777.0, posinf=1e30, neginf=-1e30)
    fb = torch.nan_to_num(b, nan=777.0, posinf=1e30, neginf=-1e30)
    return torch.allclose(fa, fb, rtol=tol, atol=tol)


def _f_for(mode, n):
    return {ops.MODE_MEDIAN: 0, ops.MODE_TRMEAN: min(3, (n - 1) // 2), ops.MODE_MEAMED: min(3, n - 1),
            ops.MODE_MEAN: 0}[mode]


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 13, 16, 17, 31, 32, 33, 63, 64, 65, 100, 128, 129])
def test_network_sorts_every_row_count(n):
    torch.manual_seed(n)
    X = torch.randn(n, 203)
    got = _host(list(X.unbind(0)), ops.MODE_MEDIAN)
    want = torch.sort(X, dim=0).values[(n - 1) // 2]
    assert torch.equal(got, want)                      # a selection, not an approximation: exact equality


def test_network_sizes_follow_merge_exchange():
    from byzpy_b200 import _C

    assert [_C.host_network_size(n) for n in (1, 2, 3, 4, 8, 16)] == [0, 1, 3, 5, 19, 63]
    assert _C.host_network_size(0) == -1 and _C.host_network_size(5000) == -1


@pytest.mark.parametrize("mode,name", MODES)
@pytest.mark.parametrize("n,d", [(2, 1), (5, 63), (8, 64), (9, 65), (10, 1000), (33, 257), (64, 4099)])
def test_matches_oracle_on_clean_data(mode, name, n, d):
    torch.manual_seed(1000 * n + d)
    rows = list(torch.randn(n, d).unbind(0))
    f = _f_for(mode, n)
    assert _same(_host(rows, mode, f), ref.cw_select(rows, mode, f))


@pytest.mark.parametrize("mode,name", MODES)
def test_matches_oracle_with_nan_inf_and_ties(mode, name):
    g = torch.Generator().manual_seed(7)
    for trial in range(40):
        n = int(torch.randint(2, 40, (1,), generator=g))
        d = int(torch.randint(1, 300, (1,), generator=g))
        X = torch.round(torch.randn(n, d, generator=g) * 2)            # heavy ties
        m = torch.rand(n, d, generator=g)
        X[m < 0.05] = float("nan")
        X[(m > 0.05) & (m < 0.08)] = float("inf")
        X[(m > 0.08) & (m < 0.10)] = float("-inf")
        rows = list(X.unbind(0))
        f = _f_for(mode, n)
        assert _same(_host(rows, mode, f, threads=1 + trial % 4), ref.cw_select(rows, mode, f)), (trial, n, d)


def test_nan_counts_as_plus_inf_and_inf_attack_rows_survive():
    rows = [torch.tensor([1.0, 2.0, 3.0]), torch.tensor([float("nan")] * 3), torch.tensor([float("inf")] * 3),
            torch.tensor([0.0, 0.0, 0.0]), torch.tensor([5.0, 5.0, 5.0])]
    assert _host(rows, ops.MODE_MEDIAN).tolist() == [5.0, 5.0, 5.0]     # sorted: 0, x, 5, inf, inf -> lower median
    assert _host(rows, ops.MODE_TRMEAN, 2).tolist() == [5.0, 5.0, 5.0]


def test_row_scales_are_applied_on_load():
    torch.manual_seed(3)
    rows = [torch.randn(777) for _ in range(9)]
    sc = [(-1.0) ** i * (1 + i) for i in range(9)]
    for mode, _ in MODES:
        f = _f_for(mode, 9)
        assert _same(_host(rows, mode, f, sc), ref.cw_select(rows, mode, f, scales=sc))


def test_thread_count_does_not_change_the_result():
    torch.manual_seed(4)
    rows = [torch.randn(64 * 700 + 5) for _ in range(11)]
    base = _host(rows, ops.MODE_MEAMED, 3, threads=1)
    for t in (2, 3, 8, 64):
        assert torch.equal(_host(rows, ops.MODE_MEAMED, 3, threads=t), base)


def test_bad_arguments_raise():
    rows = [torch.randn(10) for _ in range(4)]
    with pytest.raises(ValueError):
        _host(rows, ops.MODE_TRMEAN, 2)              # 2f < n violated
    with pytest.raises(ValueError):
        _host(rows, ops.MODE_MEAMED, 4)
    with pytest.raises(ValueError):
        _host(rows, 9)
    with pytest.raises(ValueError):
        _host(rows, ops.MODE_MEDIAN, 0, scales=[1.0])


def test_ops_dispatch_uses_the_host_kernel_for_cpu_fp32_and_falls_back_otherwise(monkeypatch):
    calls = []
    real = ops._host_cw_select

    def spy(*a, **k):
        res = real(*a, **k)
        calls.append(res is not None)
        return res

    monkeypatch.setattr(ops, "_host_cw_select", spy)
    torch.manual_seed(5)
    rows = [torch.randn(500) for _ in range(7)]
    a = ops.cw_select(rows, ops.MODE_MEDIAN)
    assert calls == [True] and torch.equal(a, ref.cw_select(rows, ops.MODE_MEDIAN))
    rows64 = [r.double() for r in rows]
    b = ops.cw_select(rows64, ops.MODE_MEDIAN)                   # fp64: PyTorch path, dtype preserved
    assert calls == [True, False] and b.dtype == torch.float64
    c = ops.cw_select(rows, ops.MODE_MEDIAN, virtual=(2, 5, 1.0, 0.5))   # synthesised rows: PyTorch path
    assert calls == [True, False, False] and c.shape == a.shape
    strided = [torch.randn(1000)[::2] for _ in range(7)]          # non-contiguous rows are compacted first
    assert torch.equal(ops.cw_select(strided, ops.MODE_MEDIAN), ref.cw_select(strided, ops.MODE_MEDIAN))
    out = torch.empty(500)
    assert ops.cw_select(rows, ops.MODE_TRMEAN, 2, out=out) is out
    assert _same(out, ref.cw_select(rows, ops.MODE_TRMEAN, 2))


def test_fused_update_still_runs_after_the_host_kernel():
    torch.manual_seed(6)
    rows = [torch.randn(300) for _ in range(5)]
    p, m = torch.zeros(300), torch.zeros(300)
    g = ops.cw_select(rows, ops.MODE_MEDIAN, update=dict(params=[p], moms=[m], lr=0.1, momentum=0.9))
    assert torch.allclose(p, -0.1 * g) and torch.allclose(m, g)


def test_aggregators_on_cpu_match_torch_formulas():
    from byzpy_b200.aggregators.coordinate_wise import (CoordinateWiseMedian, CoordinateWiseTrimmedMean,
                                                        MeanOfMedians)

    torch.manual_seed(8)
    grads = [torch.randn(4, 50) for _ in range(10)]            # shaped gradients keep their shape
    X = torch.stack([g.reshape(-1) for g in grads])
    med = CoordinateWiseMedian().aggregate(grads)
    assert med.shape == (4, 50) and torch.equal(med.reshape(-1), X.median(dim=0).values)
    tm = CoordinateWiseTrimmedMean(f=2).aggregate(grads)
    assert torch.allclose(tm.reshape(-1), X.sort(dim=0).values[2:8].mean(dim=0), atol=1e-6)
    mm = MeanOfMedians(f=3).aggregate(grads).reshape(-1)
    idx = (X - X.median(dim=0).values).abs().argsort(dim=0)[:7]
    assert torch.allclose(mm, torch.take_along_dim(X, idx, dim=0).mean(dim=0), atol=1e-6)


# ------------------------------------------------------------------------------ column statistics
@pytest.mark.parametrize("n,d", [(1, 5), (3, 257), (10, 70001), (64, 4096)])
@pytest.mark.parametrize("a,b", [(1.0, 0.0), (-1.0, 0.0), (1.0, 1.5), (0.3, -2.0)])
def test_host_colstat_matches_the_torch_formula(n, d, a, b):
    torch.manual_seed(n * 31 + d)
    rows = [torch.randn(d) for _ in range(n)]
    torch.testing.assert_close(ops.colstat(rows, a, b), ref.colstat(rows, a, b), rtol=1e-5, atol=1e-6)
    sc = [1.0 + 0.1 * i for i in range(n)]
    torch.testing.assert_close(ops.colstat(rows, a, b, scales=sc), ref.colstat(rows, a, b, scales=sc),
                               rtol=1e-5, atol=1e-6)


def test_host_colstat_propagates_non_finite_columns_like_torch_and_honours_out():
    rows = [torch.tensor([1.0, float("inf"), float("nan"), 4.0]), torch.tensor([2.0, 1.0, 1.0, 4.0])]
    got, want = ops.colstat(rows, 1.0, 1.0), ref.colstat(rows, 1.0, 1.0)
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert torch.equal(got[~torch.isnan(got)], want[~torch.isnan(want)])
    out = torch.empty(4)
    assert ops.colstat(rows, 2.0, 0.0, out=out) is out and out[0] == 3.0 and out[3] == 8.0
    dbl = [r.double() for r in rows]
    assert ops.colstat(dbl, 1.0, 0.0).dtype == torch.float64          # other dtypes stay on the torch path


def test_little_and_empire_attacks_on_cpu_use_the_same_numbers():
    from byzpy_b200.attacks import EmpireAttack, LittleAttack

    torch.manual_seed(9)
    honest = [torch.randn(3, 100) for _ in range(7)]
    X = torch.stack([h.reshape(-1) for h in honest])
    emp = EmpireAttack(scale=-2.0).apply(honest_grads=honest)
    assert emp.shape == (3, 100)
    torch.testing.assert_close(emp.reshape(-1), -2.0 * X.mean(dim=0), rtol=1e-5, atol=1e-6)
    lit = LittleAttack(f=2).apply(honest_grads=honest).reshape(-1)
    mu, sd = X.mean(dim=0), X.std(dim=0, unbiased=False)
    z = (lit - mu) / sd
    assert float(z.std()) < 1e-3                                        # one z for every coordinate
