"""Distances from an fp32-product Gram drown when the gradients share a component much larger than their
differences (with base ~ N(0, 100), d = 1e5 and 1e-4 perturbations Krum picked a wrong row
10 times out of 10).  ``center="median"`` runs the shift-invariant solvers on the rows translated by their
coordinate-wise median, which restores the ranking; the result is unchanged where the default is already exact."""
import numpy as np
import pytest
import torch

from byzpy_b200.aggregators.geometric_wise import GeometricMedian, Krum, MinimumDiameterAveraging, MoNNA, MultiKrum
from byzpy_b200.aggregators.norm_wise import CenteredClipping, ComparativeGradientElimination


def clustered(seed, n=10, d=100_000, n_out=3):
    """n - n_out rows = common + 1e-4 * noise_i (row 2 exactly at the cluster centre: the Krum winner),
    n_out rows = common + 1e-2 * noise (an outlier group 100 x further out)."""
    g = torch.Generator().manual_seed(seed)
    common = torch.randn(d, generator=g) * 100.0
    rows = []
    for i in range(n - n_out):
        eps = torch.randn(d, generator=g) * (0.0 if i == 2 else 1e-4)
        rows.append(common + eps)
    for _ in range(n_out):
        rows.append(common + torch.randn(d, generator=g) * 1e-2)
    return rows


def krum_winner_fp64(rows, f):
    X = torch.stack(rows).double()
    D = torch.cdist(X, X) ** 2
    n = len(rows)
    sc = D.sort(dim=1).values[:, 1: n - f - 1 + 1].sum(dim=1)
    return int(sc.argmin())


def test_centering_restores_the_krum_ranking_on_clustered_gradients():
    wrong_default = wrong_centered = 0
    for seed in range(6):
        rows = clustered(seed)
        want = krum_winner_fp64(rows, f=3)
        assert want == 2
        plain = Krum(f=3).aggregate(rows)
        agg = Krum(f=3)
        agg.center = "median"
        centered = agg.aggregate(rows)
        wrong_default += int(not torch.equal(plain, rows[want]))
        wrong_centered += int(not torch.equal(centered, rows[want]))
    assert wrong_centered == 0
    assert wrong_default >= 3, "the uncentred fp32 Gram was expected to lose the ranking on this input"


def test_centered_multikrum_never_selects_the_outlier_group():
    for seed in range(4):
        rows = clustered(seed)
        agg = MultiKrum(f=3, q=4)
        agg.center = "median"
        out = agg.aggregate(rows)
        X = torch.stack(rows).double()
        # the mean of 4 cluster rows is within ~1e-4 * sqrt(d) of the cluster centre; any outlier row in the mix
        # would move it by ~1e-2 * sqrt(d) / 4
        assert float((out.double() - X[2]).norm()) < 5e-4 * np.sqrt(X.shape[1])


@pytest.mark.parametrize("mk", [lambda: MultiKrum(f=2, q=3), lambda: Krum(f=2), lambda: GeometricMedian(),
                                lambda: GeometricMedian(init="mean"), lambda: MinimumDiameterAveraging(f=2),
                                lambda: MoNNA(f=2), lambda: CenteredClipping(c_tau=5.0, M=5),
                                lambda: CenteredClipping(c_tau=5.0, M=5, init="median")])
@pytest.mark.parametrize("mode", ["median", "row0"])
def test_centering_does_not_change_well_conditioned_results(mk, mode):
    g = torch.Generator().manual_seed(3)
    rows = [torch.randn(777, generator=g) + (4.0 if i >= 7 else 0.0) for i in range(9)]
    ref = mk().aggregate(rows)
    agg = mk()
    agg.center = mode
    torch.testing.assert_close(agg.aggregate(rows), ref, rtol=1e-4, atol=1e-5)


def test_centering_is_ignored_where_the_solver_needs_absolute_norms(monkeypatch):
    rows = [torch.randn(100) for _ in range(7)]
    monkeypatch.setenv("BYZPY_GRAM_CENTER", "median")
    for agg in (ComparativeGradientElimination(f=2), CenteredClipping(c_tau=1.0, init="zero")):
        assert agg._centering() is None
    assert Krum(f=2)._centering() == "median"
    monkeypatch.setenv("BYZPY_GRAM_CENTER", "off")
    assert Krum(f=2)._centering() is None
    k = Krum(f=2)
    k.center = "nonsense"
    with pytest.raises(ValueError):
        k.aggregate(rows)


def test_centering_survives_non_finite_rows():
    g = torch.Generator().manual_seed(5)
    rows = [torch.randn(500, generator=g) for _ in range(9)]
    rows[7] = torch.full((500,), float("inf"))
    rows[8][::3] = float("nan")
    agg = MultiKrum(f=2, q=3)
    agg.center = "median"
    out = agg.aggregate(rows)
    assert torch.isfinite(out).all()
    torch.testing.assert_close(out, MultiKrum(f=2, q=3).aggregate(rows), rtol=1e-4, atol=1e-5)
