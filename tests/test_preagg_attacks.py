import math
import random

import pytest
import torch

from byzpy_b200.attacks import (Attack, EmpireAttack, GaussianAttack, InfAttack, LabelFlipAttack,
                                LittleAttack, MimicAttack, SignFlipAttack)
from byzpy_b200.engine.graph.operator import OpContext
from byzpy_b200.pre_aggregators import ARC, Bucketing, Clipping, NearestNeighborMixing, PreAggregator


def vecs(n=8, d=33, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(d, generator=g) * (1 + i) for i in range(n)]


def ctx(pool=4):
    return OpContext(node_name="t", metadata={"pool_size": pool})


def test_clipping_norms_bounded_and_small_untouched():
    xs = vecs()
    out = Clipping(threshold=3.0).pre_aggregate(xs)
    for x, y in zip(xs, out):
        assert y.norm() <= 3.0 + 1e-4
        if x.norm() <= 3.0:
            assert torch.allclose(x, y)
        else:
            assert torch.allclose(y, x * (3.0 / x.norm()), atol=1e-5)


def test_arc_threshold_formula():
    xs = vecs(n=10)
    f = 3
    norms = sorted(x.norm().item() for x in xs)
    nb = int(math.floor(2.0 * f / 10 * (10 - f)))
    tau = norms[10 - nb - 1]
    out = ARC(f=f).pre_aggregate(xs)
    for x, y in zip(xs, out):
        assert y.norm().item() <= tau * (1 + 1e-5) + 1e-6
    assert all(torch.allclose(a, b) for a, b in zip(ARC(f=0).pre_aggregate(xs), xs))
    with pytest.raises(ValueError):
        ARC(f=11).pre_aggregate(xs)


def test_bucketing_means_and_permutation():
    xs = vecs(n=7)
    perm = [6, 0, 3, 2, 5, 1, 4]
    out = Bucketing(bucket_size=3, perm=perm).pre_aggregate(xs)
    assert len(out) == 3
    assert torch.allclose(out[0], (xs[6] + xs[0] + xs[3]) / 3, atol=1e-6)
    assert torch.allclose(out[2], xs[4], atol=1e-6)
    a = Bucketing(bucket_size=2, rng=random.Random(5)).pre_aggregate(xs)
    b = Bucketing(bucket_size=2, rng=random.Random(5)).pre_aggregate(xs)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    with pytest.raises(ValueError):
        Bucketing(bucket_size=2, perm=[0, 1]).pre_aggregate(xs)
    with pytest.raises(ValueError):
        Bucketing(bucket_size=0)


def test_nnm_matches_bruteforce():
    xs = vecs(n=9, seed=3)
    f = 3
    X = torch.stack(xs)
    D = torch.cdist(X, X) ** 2
    exp = [X[D[i].topk(9 - f, largest=False).indices].mean(0) for i in range(9)]
    out = NearestNeighborMixing(f=f).pre_aggregate(xs)
    assert all(torch.allclose(a, b, atol=1e-5) for a, b in zip(out, exp))
    with pytest.raises(ValueError):
        NearestNeighborMixing(f=9).pre_aggregate(xs)


@pytest.mark.parametrize("mk", [lambda: Clipping(threshold=5.0), lambda: ARC(f=2), lambda: NearestNeighborMixing(f=2)])
def test_preagg_chunked_equals_direct(mk):
    xs = vecs(n=8, d=211)
    op = mk()
    inputs = {"vectors": xs}
    tasks = list(op.create_subtasks(inputs, context=ctx()))
    assert tasks
    out = op.reduce_subtasks([t.fn(*t.args) for t in tasks], inputs, context=ctx())
    ref = mk().pre_aggregate(xs)
    assert all(torch.allclose(a, b, atol=1e-5) for a, b in zip(out, ref))
    assert isinstance(op, PreAggregator) and op.input_key == "vectors"


def test_preagg_names():
    assert Clipping().name == "pre-agg/clipping" and Bucketing(2).name == "pre-agg/bucketing"
    assert NearestNeighborMixing(1).name == "pre-agg/nnm" and ARC().name == "pre-agg/arc"


def test_sign_flip_empire_mimic_inf():
    g = vecs(n=5)
    assert torch.equal(SignFlipAttack().apply(base_grad=g[0]), -g[0])
    assert torch.allclose(EmpireAttack(scale=-2.0).apply(honest_grads=g), -2.0 * torch.stack(g).mean(0), atol=1e-6)
    assert torch.equal(MimicAttack(epsilon=3).apply(honest_grads=g), g[3])
    out = InfAttack().apply(honest_grads=g)
    assert out.shape == g[0].shape and torch.isinf(out).all() and (out > 0).all()
    with pytest.raises(ValueError):
        MimicAttack(epsilon=9).apply(honest_grads=g)
    with pytest.raises(ValueError):
        EmpireAttack().apply(honest_grads=[])


def test_little_attack_formula():
    g = vecs(n=6, seed=2)
    f = 2
    N = 6 + f
    s = max(1, N // 2 + 1 - f)
    from statistics import NormalDist

    z = NormalDist().inv_cdf((N - s) / N)
    X = torch.stack(g)
    exp = X.mean(0) + z * X.std(0, unbiased=False)
    assert torch.allclose(LittleAttack(f=f).apply(honest_grads=g), exp, atol=1e-5)
    assert LittleAttack(f=1, N=20).z_value(6) != LittleAttack(f=1).z_value(6)


def test_gaussian_reseeded_each_call():
    g = vecs(n=2, d=1000)
    a = GaussianAttack(mu=1.0, sigma=2.0, seed=3)
    x, y = a.apply(honest_grads=g), a.apply(honest_grads=g)
    assert torch.equal(x, y) and x.shape == g[0].shape
    assert abs(x.mean().item() - 1.0) < 0.3 and abs(x.std().item() - 2.0) < 0.3


def test_label_flip_gradient():
    torch.manual_seed(0)
    m = torch.nn.Linear(6, 4)
    x, y = torch.randn(5, 6), torch.tensor([0, 1, 2, 3, 0])
    out = LabelFlipAttack(num_classes=4, scale=2.0).apply(model=m, x=x, y=y)
    m.zero_grad()
    torch.nn.functional.cross_entropy(m(x), 3 - y).backward()
    exp = 2.0 * torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    assert torch.allclose(out, exp, atol=1e-6)
    m2 = torch.nn.Linear(6, 4)
    LabelFlipAttack(mapping={0: 1}).apply(model=m2, x=x, y=y)
    assert all(p.grad.abs().sum() == 0 for p in m2.parameters())
    with pytest.raises(ValueError):
        LabelFlipAttack()


def test_attack_compute_routes_inputs_by_flags():
    g = vecs(n=3)
    assert torch.equal(SignFlipAttack().compute({"base_grad": g[0], "junk": 1}, context=ctx()), -g[0])
    with pytest.raises(KeyError):
        SignFlipAttack().compute({}, context=ctx())
    with pytest.raises(KeyError):
        EmpireAttack().compute({"base_grad": g[0]}, context=ctx())
    assert issubclass(EmpireAttack, Attack) and EmpireAttack.uses_honest_grads and SignFlipAttack.uses_base_grad


@pytest.mark.parametrize("mk,key", [(lambda: EmpireAttack(scale=-1.0, chunk_size=16), "honest_grads"),
                                    (lambda: LittleAttack(f=2, chunk_size=16), "honest_grads")])
def test_attack_chunked_equals_direct(mk, key):
    g = vecs(n=6, d=100)
    op = mk()
    inputs = {key: g}
    tasks = list(op.create_subtasks(inputs, context=ctx()))
    out = op.reduce_subtasks([t.fn(*t.args) for t in tasks], inputs, context=ctx())
    assert torch.allclose(out, mk().apply(honest_grads=g), atol=1e-6)


def test_folds_for_fused_round():
    assert SignFlipAttack(scale=-3.0).fold(4).kind == "scale"
    fl = LittleAttack(f=2).fold(6)
    assert fl.kind == "virtual" and fl.a == 1.0 and fl.b == LittleAttack(f=2).z_value(6)
    assert EmpireAttack(scale=-1.5).fold(3).a == -1.5 and MimicAttack(2).fold(3).kind == "alias"
