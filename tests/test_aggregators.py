"""Operator-library unit tests: golden tensors, chunked == direct, barriered == direct, validation,
and (when the reference is installed under baseline/_ref) direct parity against it."""
import asyncio
import os
import sys

import pytest
import torch

from byzpy_b200.aggregators import Aggregator
from byzpy_b200.aggregators.coordinate_wise import (CoordinateWiseMedian, CoordinateWiseTrimmedMean,
                                                     MeanOfMedians)
from byzpy_b200.aggregators.geometric_wise import (SMEA, GeometricMedian, Krum,
                                                    MinimumDiameterAveraging, MoNNA, MultiKrum)
from byzpy_b200.aggregators.norm_wise import CAF, CenteredClipping, ComparativeGradientElimination
from byzpy_b200.engine.graph.operator import OpContext
from byzpy_b200.engine.storage.shared_store import cleanup_tensor, register_tensor

ALL = [
    lambda: CoordinateWiseMedian(chunk_size=7), lambda: CoordinateWiseTrimmedMean(f=2, chunk_size=5),
    lambda: MeanOfMedians(f=2, chunk_size=9), lambda: MultiKrum(f=2, q=3), lambda: Krum(f=2),
    lambda: GeometricMedian(), lambda: GeometricMedian(init="mean"),
    lambda: MinimumDiameterAveraging(f=2), lambda: MoNNA(f=2, reference_index=1), lambda: SMEA(f=2),
    lambda: CenteredClipping(c_tau=0.7, M=6), lambda: CenteredClipping(c_tau=0.7, init="median"),
    lambda: CenteredClipping(c_tau=2.0, init="zero"), lambda: ComparativeGradientElimination(f=2),
    lambda: CAF(f=2),
]


def grads(n=9, d=41, seed=0, shape=None):
    g = torch.Generator().manual_seed(seed)
    out = [torch.randn(d, generator=g) + (5.0 if i >= n - 2 else 0.0) for i in range(n)]
    return [x.reshape(shape) for x in out] if shape else out


def ctx(pool_size=4):
    return OpContext(node_name="t", metadata={"pool_size": pool_size})


def test_median_golden_lower_median():
    g = [torch.tensor([1.0, 10.0]), torch.tensor([2.0, 20.0]), torch.tensor([3.0, 30.0]), torch.tensor([4.0, 40.0])]
    assert torch.equal(CoordinateWiseMedian().aggregate(g), torch.tensor([2.0, 20.0]))
    assert torch.equal(CoordinateWiseMedian().aggregate(g[:3]), torch.tensor([2.0, 20.0]))


def test_trimmed_mean_golden():
    g = [torch.tensor([float(v)]) for v in (100, 1, 2, 3, -50)]
    assert CoordinateWiseTrimmedMean(f=1).aggregate(g).item() == pytest.approx(2.0)
    with pytest.raises(ValueError):
        CoordinateWiseTrimmedMean(f=3).aggregate(g)


def test_meamed_golden():
    g = [torch.tensor([float(v)]) for v in (0, 1, 2, 3, 100)]
    # median 2; the 3 closest values are 1, 2, 3
    assert MeanOfMedians(f=2).aggregate(g).item() == pytest.approx(2.0)


def test_krum_picks_cluster_member_and_multikrum_mean():
    g = [torch.zeros(4), torch.full((4,), 0.1), torch.full((4,), -0.1), torch.full((4,), 50.0)]
    out = Krum(f=1).aggregate(g)
    assert torch.equal(out, g[0])
    mk = MultiKrum(f=1, q=3).aggregate(g)
    assert torch.allclose(mk, torch.zeros(4), atol=1e-6)
    with pytest.raises(ValueError):
        MultiKrum(f=3, q=1).aggregate(g)
    with pytest.raises(ValueError):
        MultiKrum(f=1, q=4).aggregate(g)


def test_cge_monna_mda_golden():
    g = [torch.tensor([1.0, 0.0]), torch.tensor([0.0, 2.0]), torch.tensor([30.0, 0.0])]
    assert torch.allclose(ComparativeGradientElimination(f=1).aggregate(g), torch.tensor([0.5, 1.0]))
    assert torch.allclose(MoNNA(f=1, reference_index=2).aggregate(g), torch.tensor([15.5, 0.0]))
    assert torch.allclose(MinimumDiameterAveraging(f=1).aggregate(g), torch.tensor([0.5, 1.0]))


def test_geometric_median_of_collinear_points_is_the_middle_one():
    g = [torch.tensor([0.0, 0.0]), torch.tensor([1.0, 0.0]), torch.tensor([10.0, 0.0])]
    out = GeometricMedian(tol=1e-9, max_iter=500).aggregate(g)
    assert torch.allclose(out, torch.tensor([1.0, 0.0]), atol=1e-3)


def test_centered_clipping_zero_radius_returns_init():
    g = grads()
    out = CenteredClipping(c_tau=0.0, M=5).aggregate(g)
    assert torch.allclose(out, torch.stack(g).mean(0), atol=1e-6)


@pytest.mark.parametrize("mk", ALL)
def test_shape_dtype_preserved_and_handles_accepted(mk):
    g = grads(shape=(41,))
    g2 = [x.reshape(1, 41).double() for x in g]
    out = mk().aggregate(g2)
    assert out.shape == (1, 41) and out.dtype == torch.float64
    handles = [register_tensor(x.numpy()) for x in g]
    try:
        via_handles = mk().aggregate(handles)
        via_dicts = mk().aggregate([{"name": h.name, "shape": h.shape, "dtype": h.dtype} for h in handles])
    finally:
        for h in handles:
            cleanup_tensor(h)
    ref = mk().aggregate(g)
    assert torch.allclose(via_handles, ref, atol=1e-6) and torch.allclose(via_dicts, ref, atol=1e-6)


@pytest.mark.parametrize("mk", ALL)
def test_chunked_equals_direct(mk):
    g = grads(n=9, d=103, seed=3)
    op = mk()
    inputs = {"gradients": g}
    tasks = list(op.create_subtasks(inputs, context=ctx()))
    assert len(tasks) >= 1
    partials = [t.fn(*t.args, **dict(t.kwargs)) for t in tasks]
    chunked = op.reduce_subtasks(partials, inputs, context=ctx())
    assert torch.allclose(chunked, mk().aggregate(g), rtol=1e-5, atol=1e-6)


def test_subtask_count_grows_with_pool_size():
    from byzpy_b200.aggregators.base import _release_packed

    g = grads(n=4, d=65536)
    counts = []
    for pool in (2, 6):
        op = CoordinateWiseMedian(chunk_size=8192)
        inputs = {"gradients": g}
        counts.append(len(list(op.create_subtasks(inputs, context=ctx(pool)))))
        _release_packed(op, inputs)
    assert counts[1] > counts[0] >= 1


class _StubPool:
    """Only what a barriered operator needs: ``size`` and ``run_subtask`` (reference test technique)."""

    size = 3

    async def run_subtask(self, st):
        return st.fn(*st.args, **dict(st.kwargs))


@pytest.mark.parametrize("mk", [lambda: GeometricMedian(), lambda: GeometricMedian(init="mean"),
                                lambda: CenteredClipping(c_tau=0.5, M=7), lambda: CenteredClipping(c_tau=0.5, init="median")])
def test_barriered_equals_direct(mk):
    g = grads(n=10, d=333, seed=5)
    op = mk()
    assert op.supports_barriered_subtasks
    out = asyncio.run(op.run({"gradients": g}, context=ctx(3), pool=_StubPool()))
    assert torch.allclose(out, mk().aggregate(g), rtol=1e-5, atol=1e-6)


def test_operator_is_reentrant_across_concurrent_invocations():
    op = CoordinateWiseMedian(chunk_size=16)
    a, b = grads(seed=1), grads(seed=2)

    async def both():
        return await asyncio.gather(op.run({"gradients": a}, context=ctx(3), pool=_StubPool2()),
                                    op.run({"gradients": b}, context=ctx(3), pool=_StubPool2()))

    ra, rb = asyncio.run(both())
    assert torch.equal(ra, torch.stack(a).median(0).values) and torch.equal(rb, torch.stack(b).median(0).values)


class _StubPool2(_StubPool):
    async def run_subtask(self, st):
        await asyncio.sleep(0)
        return st.fn(*st.args, **dict(st.kwargs))


def test_compute_contract_errors():
    op = CoordinateWiseMedian()
    with pytest.raises(KeyError):
        op.compute({}, context=ctx())
    with pytest.raises(TypeError):
        op.compute({"gradients": 3}, context=ctx())
    with pytest.raises(ValueError):
        op.aggregate([])
    assert isinstance(op, Aggregator) and op.name == "coordinate-wise-median" and op.input_key == "gradients"


def test_names_match_reference_contract():
    names = {type(mk()).__name__: mk().name for mk in ALL}
    assert names["CoordinateWiseTrimmedMean"] == "coordinate-wise-trimmed-mean"
    assert names["MeanOfMedians"] == "mean-of-medians" and names["MultiKrum"] == "multi-krum"
    assert names["Krum"] == "krum" and names["GeometricMedian"] == "geometric-median"
    assert names["MinimumDiameterAveraging"] == "minimum-diameter-averaging"
    assert names["MoNNA"] == "monna" and names["SMEA"] == "smea" and names["CAF"] == "caf"
    assert names["CenteredClipping"] == "centered-clipping"
    assert names["ComparativeGradientElimination"] == "comparative-gradient-elimination"


def test_inf_attack_rows_do_not_poison_robust_aggregators():
    g = grads(n=9)
    g[-1] = torch.full_like(g[0], float("inf"))
    for mk in [lambda: CoordinateWiseMedian(), lambda: CoordinateWiseTrimmedMean(f=2), lambda: MultiKrum(f=2, q=3),
               lambda: ComparativeGradientElimination(f=2), lambda: MoNNA(f=2)]:
        assert torch.isfinite(mk().aggregate(g)).all()


REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "byzpy")), reason="reference not installed")
def test_parity_with_reference_direct_paths():
    sys.path.insert(0, REF)
    try:
        import byzpy.aggregators.coordinate_wise as rcw
        import byzpy.aggregators.geometric_wise as rgw
        import byzpy.aggregators.norm_wise as rnw
    finally:
        sys.path.remove(REF)
    g = grads(n=11, d=57, seed=9)
    cases = [(CoordinateWiseMedian, rcw.CoordinateWiseMedian, {}), (CoordinateWiseTrimmedMean, rcw.CoordinateWiseTrimmedMean, {"f": 3}),
             (MeanOfMedians, rcw.MeanOfMedians, {"f": 3}), (MultiKrum, rgw.MultiKrum, {"f": 3, "q": 4}),
             (Krum, rgw.Krum, {"f": 3}), (GeometricMedian, rgw.GeometricMedian, {}),
             (MinimumDiameterAveraging, rgw.MinimumDiameterAveraging, {"f": 3}), (MoNNA, rgw.MoNNA, {"f": 3}),
             (SMEA, rgw.SMEA, {"f": 3}), (CenteredClipping, rnw.CenteredClipping, {"c_tau": 1.3}),
             (ComparativeGradientElimination, rnw.ComparativeGradientElimination, {"f": 3}), (CAF, rnw.CAF, {"f": 3})]
    for mine, theirs, kw in cases:
        assert torch.allclose(mine(**kw).aggregate(g), theirs(**kw).aggregate(g), rtol=1e-4, atol=1e-5), mine.__name__
