"""Operator library tiers: every aggregator / pre-aggregator / attack against an independent
brute-force oracle written in plain NumPy on random data, plus the algebraic invariances the
definitions imply (permutation, translation, scaling), degenerate inputs and argument validation
(mirrors the per-operator files under reference tests/aggregators, tests/pre_aggregators,
tests/attacks)."""
import itertools
import math

import numpy as np
import pytest
import torch

from byzpy_b200.aggregators.coordinate_wise import (CoordinateWiseMedian, CoordinateWiseTrimmedMean,
                                                     MeanOfMedians)
from byzpy_b200.aggregators.geometric_wise import (SMEA, GeometricMedian, Krum,
                                                    MinimumDiameterAveraging, MoNNA, MultiKrum)
from byzpy_b200.aggregators.norm_wise import CAF, CenteredClipping, ComparativeGradientElimination
from byzpy_b200.attacks import (EmpireAttack, GaussianAttack, InfAttack, LabelFlipAttack, LittleAttack,
                                MimicAttack, SignFlipAttack)
from byzpy_b200.engine.graph.operator import OpContext
from byzpy_b200.pre_aggregators import ARC, Bucketing, Clipping, NearestNeighborMixing


def data(n, d, seed=0, outliers=0, dtype=torch.float32):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d))
    if outliers:
        X[n - outliers:] += 8.0
    return X, [torch.tensor(row, dtype=dtype) for row in X]


def close(t, ref, tol=1e-5):
    return np.allclose(t.double().numpy(), ref, rtol=tol, atol=tol)


SHAPES = [(5, 7, 0), (8, 33, 1), (11, 64, 2), (6, 1, 3), (16, 19, 4)]


# ------------------------------------------------------------------------------------ NumPy oracles
def o_median(X):
    return np.sort(X, axis=0)[(len(X) - 1) // 2]


def o_trimmed(X, f):
    S = np.sort(X, axis=0)
    return S[f:len(X) - f].mean(axis=0)


def o_meamed(X, f):
    med = o_median(X)
    out = np.empty(X.shape[1])
    for j in range(X.shape[1]):
        order = np.argsort(np.abs(X[:, j] - med[j]), kind="stable")
        out[j] = X[order[:len(X) - f], j].mean()
    return out


def o_sqdist(X):
    return ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)


def o_multikrum(X, f, q):
    D = o_sqdist(X)
    n = len(X)
    scores = np.array([np.sort(np.delete(D[i], i))[:n - f - 1].sum() for i in range(n)])
    return X[np.argsort(scores, kind="stable")[:q]].mean(axis=0)


def o_mda(X, f):
    D = o_sqdist(X)
    best = min(itertools.combinations(range(len(X)), len(X) - f),
               key=lambda s: max(D[i, j] for i in s for j in s))
    return X[list(best)].mean(axis=0)


def o_smea(X, f):
    def top_eig(s):
        Y = X[list(s)] - X[list(s)].mean(axis=0)
        return np.linalg.eigvalsh(Y @ Y.T / len(s))[-1]

    best = min(itertools.combinations(range(len(X)), len(X) - f), key=top_eig)
    return X[list(best)].mean(axis=0)


def o_monna(X, f, ref):
    d = ((X - X[ref]) ** 2).sum(-1)
    return X[np.argsort(d, kind="stable")[:len(X) - f]].mean(axis=0)


def o_cge(X, f):
    return X[np.argsort((X ** 2).sum(-1), kind="stable")[:len(X) - f]].mean(axis=0)


def o_centered_clip(X, c_tau, M, v0):
    v = v0.copy()
    for _ in range(M):
        diff = X - v
        norms = np.linalg.norm(diff, axis=1)
        scale = np.minimum(1.0, c_tau / np.maximum(norms, 1e-12))
        v = v + (diff * scale[:, None]).mean(axis=0)
    return v


def o_weiszfeld(X, iters=2000):
    z = X.mean(axis=0)
    for _ in range(iters):
        w = 1.0 / np.maximum(np.linalg.norm(X - z, axis=1), 1e-12)
        z = (w[:, None] * X).sum(0) / w.sum()
    return z


# ------------------------------------------------------------------------ aggregators versus oracles
@pytest.mark.parametrize("n,d,seed", SHAPES)
def test_median_oracle(n, d, seed):
    X, vs = data(n, d, seed)
    assert close(CoordinateWiseMedian().aggregate(vs), o_median(X))


@pytest.mark.parametrize("n,d,seed", SHAPES)
@pytest.mark.parametrize("f", [0, 1, 2])
def test_trimmed_mean_oracle(n, d, seed, f):
    X, vs = data(n, d, seed)
    assert close(CoordinateWiseTrimmedMean(f=f).aggregate(vs), o_trimmed(X, f))


@pytest.mark.parametrize("n,d,seed", SHAPES)
@pytest.mark.parametrize("f", [0, 1, 3])
def test_meamed_oracle(n, d, seed, f):
    X, vs = data(n, d, seed)
    assert close(MeanOfMedians(f=f).aggregate(vs), o_meamed(X, f))


@pytest.mark.parametrize("n,d,seed", SHAPES)
@pytest.mark.parametrize("f,q", [(0, 1), (1, 2), (2, 3), (1, 1)])
def test_multikrum_oracle(n, d, seed, f, q):
    X, vs = data(n, d, seed, outliers=f)
    assert close(MultiKrum(f=f, q=q).aggregate(vs), o_multikrum(X, f, q))
    if q == 1:
        assert close(Krum(f=f).aggregate(vs), o_multikrum(X, f, 1))


@pytest.mark.parametrize("n,d,seed", [(5, 7, 0), (8, 33, 1), (9, 12, 2), (6, 1, 3)])
@pytest.mark.parametrize("f", [0, 1, 2])
def test_mda_oracle(n, d, seed, f):
    X, vs = data(n, d, seed, outliers=f)
    assert close(MinimumDiameterAveraging(f=f).aggregate(vs), o_mda(X, f))


@pytest.mark.parametrize("n,d,seed", [(5, 7, 0), (8, 33, 1), (9, 12, 2)])
@pytest.mark.parametrize("f", [0, 1, 2])
def test_smea_oracle(n, d, seed, f):
    X, vs = data(n, d, seed, outliers=f)
    assert close(SMEA(f=f).aggregate(vs), o_smea(X, f), tol=1e-4)


@pytest.mark.parametrize("n,d,seed", SHAPES)
@pytest.mark.parametrize("f,ref", [(0, 0), (1, 2), (2, 4)])
def test_monna_oracle(n, d, seed, f, ref):
    X, vs = data(n, d, seed, outliers=f)
    assert close(MoNNA(f=f, reference_index=ref).aggregate(vs), o_monna(X, f, ref))


@pytest.mark.parametrize("n,d,seed", SHAPES)
@pytest.mark.parametrize("f", [0, 1, 4])
def test_cge_oracle(n, d, seed, f):
    X, vs = data(n, d, seed, outliers=f)
    assert close(ComparativeGradientElimination(f=f).aggregate(vs), o_cge(X, f))


@pytest.mark.parametrize("n,d,seed", SHAPES)
@pytest.mark.parametrize("c_tau,M,init", [(0.5, 1, "mean"), (1.5, 10, "mean"), (0.7, 4, "zero"), (0.3, 7, "median"),
                                          (100.0, 3, "zero")])
def test_centered_clipping_oracle(n, d, seed, c_tau, M, init):
    X, vs = data(n, d, seed, outliers=1)
    v0 = {"mean": X.mean(0), "zero": np.zeros(d), "median": o_median(X)}[init]
    assert close(CenteredClipping(c_tau=c_tau, M=M, init=init).aggregate(vs), o_centered_clip(X, c_tau, M, v0), tol=1e-4)


@pytest.mark.parametrize("n,d,seed", [(5, 7, 0), (8, 33, 1), (11, 64, 2), (16, 19, 4)])
@pytest.mark.parametrize("init", ["median", "mean"])
def test_geometric_median_is_the_fermat_point(n, d, seed, init):
    X, vs = data(n, d, seed, outliers=2)
    z = GeometricMedian(init=init, tol=1e-10, max_iter=2000).aggregate(vs).double().numpy()
    unit = (X - z) / np.linalg.norm(X - z, axis=1, keepdims=True)
    assert np.linalg.norm(unit.sum(0)) < 1e-3                       # first-order optimality
    assert np.allclose(z, o_weiszfeld(X), atol=1e-3)
    cost = lambda p: np.linalg.norm(X - p, axis=1).sum()
    assert cost(z) <= cost(X.mean(0)) + 1e-6 and cost(z) <= cost(o_median(X)) + 1e-6


def test_geometric_median_max_iter_zero_returns_start_point():
    X, vs = data(7, 9, 5)
    assert close(GeometricMedian(init="mean", max_iter=0).aggregate(vs), X.mean(0))
    assert close(GeometricMedian(init="median", max_iter=0).aggregate(vs), o_median(X))


def test_centered_clipping_large_radius_one_round_reaches_the_mean():
    X, vs = data(9, 5, 6)
    assert close(CenteredClipping(c_tau=1e9, M=1, init="zero").aggregate(vs), X.mean(0))
    assert close(CenteredClipping(c_tau=1.0, M=0, init="mean").aggregate(vs), X.mean(0))


def test_caf_filters_a_far_cluster_and_matches_mean_without_outliers():
    X, vs = data(12, 20, 7, outliers=3)
    out = CAF(f=3).aggregate(vs).double().numpy()
    honest = X[:9].mean(0)
    assert np.linalg.norm(out - honest) < 0.25 * np.linalg.norm(X.mean(0) - honest)
    Xc, vc = data(10, 6, 8)
    assert np.linalg.norm(CAF(f=0).aggregate(vc).double().numpy() - Xc.mean(0)) < 1e-4


# ----------------------------------------------------------------------------------- invariances
def _fam():
    return {
        "median": CoordinateWiseMedian(), "trmean": CoordinateWiseTrimmedMean(f=2), "meamed": MeanOfMedians(f=2),
        "multikrum": MultiKrum(f=2, q=3), "krum": Krum(f=2), "gm": GeometricMedian(tol=1e-9, max_iter=1000),
        "mda": MinimumDiameterAveraging(f=2), "smea": SMEA(f=2), "monna": MoNNA(f=2),
        "cclip": CenteredClipping(c_tau=0.8, M=5), "cge": ComparativeGradientElimination(f=2), "caf": CAF(f=2),
    }


@pytest.mark.parametrize("name", [k for k in _fam() if k != "monna"])
def test_permutation_invariance(name):
    _, vs = data(9, 23, 11, outliers=2)
    perm = [4, 0, 8, 2, 6, 1, 7, 3, 5]
    a, b = _fam()[name].aggregate(vs), _fam()[name].aggregate([vs[i] for i in perm])
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", [k for k in _fam() if k not in ("cge",)])
def test_translation_equivariance(name):
    _, vs = data(9, 23, 12, outliers=2)
    shift = torch.linspace(-3, 3, 23)
    a = _fam()[name].aggregate(vs) + shift
    b = _fam()[name].aggregate([v + shift for v in vs])
    assert torch.allclose(a, b, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("name", [k for k in _fam() if k not in ("cclip",)])
@pytest.mark.parametrize("scale", [0.25, 3.0])
def test_positive_scaling_equivariance(name, scale):
    _, vs = data(9, 23, 13, outliers=2)
    a = _fam()[name].aggregate(vs) * scale
    b = _fam()[name].aggregate([v * scale for v in vs])
    assert torch.allclose(a, b, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("name", list(_fam()))
def test_identical_inputs_return_that_vector(name):
    v = torch.linspace(-1, 1, 17)
    out = _fam()[name].aggregate([v.clone() for _ in range(7)])
    assert torch.allclose(out, v, atol=1e-5)


@pytest.mark.parametrize("name", list(_fam()))
def test_output_lies_in_the_coordinate_box_of_the_inputs(name):
    _, vs = data(9, 31, 14, outliers=2)
    S = torch.stack(vs)
    out = _fam()[name].aggregate(vs)
    assert (out <= S.max(0).values + 1e-4).all() and (out >= S.min(0).values - 1e-4).all()


@pytest.mark.parametrize("name", list(_fam()))
def test_float64_inputs_keep_dtype_and_agree_with_float32(name):
    _, v64 = data(9, 13, 15, outliers=2, dtype=torch.float64)
    out64 = _fam()[name].aggregate(v64)
    out32 = _fam()[name].aggregate([v.float() for v in v64])
    assert out64.dtype == torch.float64 and out32.dtype == torch.float32
    assert torch.allclose(out64.float(), out32, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("name", list(_fam()))
def test_inputs_are_not_modified(name):
    _, vs = data(9, 13, 16, outliers=2)
    before = [v.clone() for v in vs]
    _fam()[name].aggregate(vs)
    assert all(torch.equal(a, b) for a, b in zip(vs, before))


@pytest.mark.parametrize("name", list(_fam()))
def test_tuple_and_matrix_shaped_inputs(name):
    _, vs = data(9, 12, 17, outliers=2)
    flat = _fam()[name].aggregate(vs)
    assert torch.allclose(_fam()[name].aggregate(tuple(vs)), flat)
    shaped = _fam()[name].aggregate([v.reshape(3, 4) for v in vs])
    assert shaped.shape == (3, 4) and torch.allclose(shaped.reshape(-1), flat)


# ---------------------------------------------------------------------- degenerate sizes + validation
def test_single_gradient_inputs():
    v = [torch.tensor([1.0, -2.0, 3.0])]
    for agg in (CoordinateWiseMedian(), CoordinateWiseTrimmedMean(f=0), MeanOfMedians(f=0), GeometricMedian(),
                MinimumDiameterAveraging(f=0), MoNNA(f=0), SMEA(f=0), CenteredClipping(c_tau=1.0),
                ComparativeGradientElimination(f=0), CAF(f=0)):
        assert torch.allclose(agg.aggregate(v), v[0], atol=1e-6), agg.name
    with pytest.raises(ValueError):
        Krum(f=0).aggregate(v)                                      # needs n >= f + 2


@pytest.mark.parametrize("mk", [lambda: CoordinateWiseTrimmedMean(f=-1), lambda: MeanOfMedians(f=-1),
                                lambda: MultiKrum(f=-1, q=1), lambda: MultiKrum(f=0, q=0), lambda: Krum(f=-1),
                                lambda: MinimumDiameterAveraging(f=-1), lambda: MoNNA(f=-1),
                                lambda: MoNNA(f=0, reference_index=-1), lambda: SMEA(f=-1), lambda: CAF(f=-1),
                                lambda: CAF(f=0, power_iters=-1), lambda: ComparativeGradientElimination(f=-1),
                                lambda: CenteredClipping(c_tau=-1.0), lambda: CenteredClipping(c_tau=1.0, M=-1),
                                lambda: CenteredClipping(c_tau=1.0, eps=0.0), lambda: CenteredClipping(c_tau=1.0, init="x"),
                                lambda: GeometricMedian(tol=0.0), lambda: GeometricMedian(max_iter=-1),
                                lambda: GeometricMedian(eps=0.0), lambda: GeometricMedian(init="zero"),
                                lambda: CoordinateWiseMedian(chunk_size=0), lambda: MultiKrum(f=0, q=1, chunk_size=0)])
def test_constructor_validation(mk):
    with pytest.raises(ValueError):
        mk()


@pytest.mark.parametrize("agg", [CoordinateWiseTrimmedMean(f=3), MeanOfMedians(f=6), MultiKrum(f=5, q=1),
                                 MultiKrum(f=1, q=6), Krum(f=5), MinimumDiameterAveraging(f=6), MoNNA(f=3),
                                 MoNNA(f=0, reference_index=6), SMEA(f=3), CAF(f=3), ComparativeGradientElimination(f=6)])
def test_f_too_large_for_n_is_rejected_at_call_time(agg):
    _, vs = data(6, 4, 18)
    with pytest.raises(ValueError):
        agg.aggregate(vs)


@pytest.mark.parametrize("name", list(_fam()))
def test_empty_and_ragged_inputs_are_rejected(name):
    with pytest.raises(ValueError):
        _fam()[name].aggregate([])
    _, vs = data(9, 8, 19)
    with pytest.raises(ValueError):
        _fam()[name].aggregate(vs[:-1] + [torch.zeros(5)])


@pytest.mark.parametrize("name", list(_fam()))
def test_compute_contract(name):
    agg = _fam()[name]
    _, vs = data(9, 8, 20)
    ctx = OpContext("n")
    assert torch.allclose(agg.compute({"gradients": vs}, context=ctx), agg.aggregate(vs))
    with pytest.raises(KeyError):
        agg.compute({"vectors": vs}, context=ctx)
    with pytest.raises(TypeError):
        agg.compute({"gradients": 3}, context=ctx)


# ---------------------------------------------------------------------------------- pre-aggregators
@pytest.mark.parametrize("n,d,seed", SHAPES)
@pytest.mark.parametrize("tau", [0.0, 0.5, 3.0, 1e6])
def test_clipping_oracle(n, d, seed, tau):
    X, vs = data(n, d, seed)
    norms = np.linalg.norm(X, axis=1, keepdims=True)
    exp = X * np.minimum(1.0, tau / np.maximum(norms, 1e-30))
    out = Clipping(threshold=tau).pre_aggregate(vs)
    assert len(out) == n and all(close(o, e) for o, e in zip(out, exp))


@pytest.mark.parametrize("n,d,seed", SHAPES)
@pytest.mark.parametrize("f", [0, 1, 2])
def test_nnm_oracle(n, d, seed, f):
    X, vs = data(n, d, seed, outliers=f)
    D = o_sqdist(X)
    exp = np.stack([X[np.argsort(D[i], kind="stable")[:n - f]].mean(0) for i in range(n)])
    out = NearestNeighborMixing(f=f).pre_aggregate(vs)
    assert len(out) == n and all(close(o, e) for o, e in zip(out, exp))


@pytest.mark.parametrize("n,d,seed", SHAPES)
@pytest.mark.parametrize("f", [0, 1, 2])
def test_arc_oracle(n, d, seed, f):
    X, vs = data(n, d, seed, outliers=1)
    norms = np.linalg.norm(X, axis=1)
    k = int(math.floor(2.0 * (f / n) * (n - f)))
    exp = X.copy()
    if k > 0:
        order = np.argsort(norms, kind="stable")
        tau = norms[order[n - k - 1]]
        for i in order[n - k:]:
            exp[i] = X[i] * (tau / norms[i])
    out = ARC(f=f).pre_aggregate(vs)
    assert all(close(o, e) for o, e in zip(out, exp))


@pytest.mark.parametrize("n,s", [(6, 2), (7, 3), (5, 1), (4, 8), (9, 4)])
def test_bucketing_oracle(n, s):
    X, vs = data(n, 10, n + s)
    perm = list(np.random.default_rng(n).permutation(n))
    out = Bucketing(bucket_size=s, perm=perm).pre_aggregate(vs)
    exp = [X[perm[i:i + s]].mean(0) for i in range(0, n, s)]
    assert len(out) == math.ceil(n / s) and all(close(o, e) for o, e in zip(out, exp))


def test_bucketing_random_order_is_seedable_and_preserves_the_grand_mean():
    import random

    X, vs = data(8, 6, 21)
    a = Bucketing(bucket_size=2, rng=random.Random(3)).pre_aggregate(vs)
    b = Bucketing(bucket_size=2, rng=random.Random(3)).pre_aggregate(vs)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert close(torch.stack(a).mean(0), X.mean(0))


@pytest.mark.parametrize("mk", [lambda: Clipping(threshold=-1.0), lambda: Clipping(chunk_size=0), lambda: ARC(f=-1),
                                lambda: ARC(chunk_size=0), lambda: Bucketing(bucket_size=0),
                                lambda: Bucketing(bucket_size=2, feature_chunk_size=0),
                                lambda: NearestNeighborMixing(f=-1), lambda: NearestNeighborMixing(f=1, feature_chunk_size=0)])
def test_preaggregator_constructor_validation(mk):
    with pytest.raises(ValueError):
        mk()


def test_preaggregator_call_time_validation_and_contract():
    _, vs = data(5, 4, 22)
    with pytest.raises(ValueError):
        NearestNeighborMixing(f=5).pre_aggregate(vs)
    with pytest.raises(ValueError):
        ARC(f=6).pre_aggregate(vs)
    with pytest.raises(ValueError):
        Bucketing(bucket_size=2, perm=[0, 1, 2, 3, 3]).pre_aggregate(vs)
    with pytest.raises(ValueError):
        Bucketing(bucket_size=2, perm=[0, 1]).pre_aggregate(vs)
    for pre in (Clipping(), ARC(f=1), Bucketing(bucket_size=2), NearestNeighborMixing(f=1)):
        with pytest.raises(ValueError):
            pre.pre_aggregate([])
        with pytest.raises(KeyError):
            pre.compute({"gradients": vs}, context=OpContext("n"))
        with pytest.raises(TypeError):
            pre.compute({"vectors": 1.0}, context=OpContext("n"))
        out = pre.compute({"vectors": vs}, context=OpContext("n"))
        assert isinstance(out, list) and all(o.shape == vs[0].shape and o.dtype == vs[0].dtype for o in out)


def test_preaggregators_keep_shape_of_matrix_inputs():
    _, vs = data(6, 12, 23)
    shaped = [v.reshape(3, 4) for v in vs]
    for pre in (Clipping(threshold=1.0), ARC(f=1), Bucketing(bucket_size=2, perm=range(6)), NearestNeighborMixing(f=1)):
        a, b = pre.pre_aggregate(shaped), pre.pre_aggregate(vs)
        assert all(x.shape == (3, 4) and torch.allclose(x.reshape(-1), y) for x, y in zip(a, b)), pre.name


# ------------------------------------------------------------------------------------------- attacks
@pytest.mark.parametrize("n,d,seed", SHAPES)
def test_column_statistic_attacks_oracle(n, d, seed):
    X, vs = data(n, d, seed)
    mu, sd = X.mean(0), X.std(0, ddof=0)
    assert close(EmpireAttack(scale=-1.5).apply(honest_grads=vs), -1.5 * mu)
    f, N = 2, n + 2
    s = math.floor(N / 2 + 1) - f
    from statistics import NormalDist

    z = NormalDist().inv_cdf(min(max((N - s) / N, 0.0), 1.0)) if 0 < (N - s) / N < 1 else 0.0
    out = LittleAttack(f=f, N=N).apply(honest_grads=vs).double().numpy()
    ratio = (out - mu) / np.where(sd > 0, sd, 1.0)
    assert np.allclose(ratio, ratio[0], atol=1e-4)                  # mu + c * sigma with one scalar c
    assert abs(abs(ratio[0]) - abs(z)) < 1e-3 or n == 1


def test_sign_flip_mimic_inf_gaussian_basics():
    _, vs = data(5, 9, 24)
    assert torch.equal(SignFlipAttack(scale=-2.0).apply(base_grad=vs[0]), -2.0 * vs[0])
    assert torch.equal(MimicAttack(epsilon=3).apply(honest_grads=vs), vs[3])
    assert MimicAttack(epsilon=3).apply(honest_grads=vs) is not vs[3]
    inf = InfAttack().apply(honest_grads=vs)
    assert inf.shape == vs[0].shape and torch.isinf(inf).all() and (inf > 0).all()
    g = GaussianAttack(mu=2.0, sigma=0.0, seed=1).apply(honest_grads=vs)
    assert torch.equal(g, torch.full_like(vs[0], 2.0))
    big = GaussianAttack(mu=1.0, sigma=3.0, seed=5).apply(honest_grads=[torch.zeros(20000)])
    assert abs(big.mean().item() - 1.0) < 0.1 and abs(big.std().item() - 3.0) < 0.1


@pytest.mark.parametrize("mk", [lambda: EmpireAttack(chunk_size=0), lambda: LittleAttack(f=-1), lambda: LittleAttack(f=1, N=0),
                                lambda: LittleAttack(f=1, chunk_size=0), lambda: GaussianAttack(sigma=-1.0),
                                lambda: GaussianAttack(chunk_size=0), lambda: InfAttack(chunk_size=0),
                                lambda: MimicAttack(epsilon=-1), lambda: MimicAttack(chunk_size=0),
                                lambda: SignFlipAttack(chunk_size=0), lambda: LabelFlipAttack()])
def test_attack_constructor_validation(mk):
    with pytest.raises(ValueError):
        mk()


def test_attack_call_time_validation():
    _, vs = data(4, 3, 25)
    for atk in (EmpireAttack(), LittleAttack(f=1), GaussianAttack(), InfAttack(), MimicAttack()):
        with pytest.raises(ValueError):
            atk.apply(honest_grads=None)
        with pytest.raises(ValueError):
            atk.apply(honest_grads=[])
    with pytest.raises(ValueError):
        SignFlipAttack().apply(base_grad=None)
    with pytest.raises(ValueError):
        MimicAttack(epsilon=4).apply(honest_grads=vs)
    with pytest.raises(ValueError):
        LittleAttack(f=9, N=5).apply(honest_grads=vs)
    with pytest.raises(ValueError):
        LabelFlipAttack(num_classes=3).apply(model=None, x=None, y=None)


def test_label_flip_mapping_and_involution():
    atk = LabelFlipAttack(mapping={0: 2, 2: 0})
    y = torch.tensor([0, 1, 2, 3, 0])
    assert atk.corrupt(y).tolist() == [2, 1, 0, 3, 2]
    inv = LabelFlipAttack(num_classes=10)
    assert inv.corrupt(torch.arange(10)).tolist() == list(range(9, -1, -1))
    assert torch.equal(inv.corrupt(inv.corrupt(y)), y)


def test_label_flip_leaves_model_grads_clean_and_scales():
    torch.manual_seed(0)
    model = torch.nn.Linear(4, 3)
    x, y = torch.randn(8, 4), torch.randint(0, 3, (8,))
    g1 = LabelFlipAttack(num_classes=3).apply(model=model, x=x, y=y)
    assert all(p.grad is None or not p.grad.any() for p in model.parameters())
    g2 = LabelFlipAttack(num_classes=3, scale=-3.0).apply(model=model, x=x, y=y)
    assert torch.allclose(g2, -3.0 * g1, atol=1e-6) and g1.numel() == sum(p.numel() for p in model.parameters())
    loss = torch.nn.functional.cross_entropy(model(x), 2 - y)
    exp = torch.cat([g.reshape(-1) for g in torch.autograd.grad(loss, list(model.parameters()))])
    assert torch.allclose(g1, exp, atol=1e-6)
