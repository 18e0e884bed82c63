"""The pool (subtask / barriered) path of EVERY operator gives the direct call's result.

``Operator.run`` normally measures and picks the faster route; ``BYZPY_POOL_DISPATCH=reference`` pins the
reference's rule (pool present -> subtasks), so these tests really go through ``create_subtasks`` /
``reduce_subtasks`` / ``run_barriered_subtasks`` -- once on thread workers (views of the caller's rows) and once on
process workers (rows packed into shared memory, results shipped back)."""
import asyncio
import glob
import random

import pytest
import torch

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean, MeanOfMedians
from byzpy_b200.aggregators.geometric_wise import (GeometricMedian, Krum, MinimumDiameterAveraging, MoNNA, MultiKrum,
                                                  SMEA)
from byzpy_b200.aggregators.norm_wise import CAF, CenteredClipping, ComparativeGradientElimination
from byzpy_b200.attacks import (EmpireAttack, GaussianAttack, InfAttack, LittleAttack, MimicAttack, SignFlipAttack)
from byzpy_b200.engine.graph.ops import make_single_operator_graph
from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig
from byzpy_b200.engine.graph.scheduler import NodeScheduler
from byzpy_b200.pre_aggregators import ARC, Bucketing, Clipping, NearestNeighborMixing

N, D = 10, 3001          # d is odd and not a multiple of any chunk size: tails everywhere


def _data():
    g = torch.Generator().manual_seed(5)
    rows = [torch.randn(D, generator=g) for _ in range(N)]
    rows[3] = rows[3] * 25.0                     # an outlier so that selections are not ties
    return rows


OPERATORS = {
    "median": (lambda: CoordinateWiseMedian(chunk_size=512), "gradients"),
    "trimmed-mean": (lambda: CoordinateWiseTrimmedMean(f=2, chunk_size=512), "gradients"),
    "meamed": (lambda: MeanOfMedians(f=2, chunk_size=512), "gradients"),
    "krum": (lambda: Krum(f=2, chunk_size=3), "gradients"),
    "multi-krum": (lambda: MultiKrum(f=2, q=4, chunk_size=3), "gradients"),
    "geometric-median": (lambda: GeometricMedian(chunk_size=3), "gradients"),
    "geometric-median-mean-init": (lambda: GeometricMedian(init="mean", chunk_size=3), "gradients"),
    "mda": (lambda: MinimumDiameterAveraging(f=2, chunk_size=7), "gradients"),
    "smea": (lambda: SMEA(f=2, chunk_size=7), "gradients"),
    "monna": (lambda: MoNNA(f=2, reference_index=1, chunk_size=3), "gradients"),
    "centered-clipping": (lambda: CenteredClipping(c_tau=0.7, M=4, chunk_size=3), "gradients"),
    "centered-clipping-median": (lambda: CenteredClipping(c_tau=0.7, M=4, init="median", chunk_size=3), "gradients"),
    "cge": (lambda: ComparativeGradientElimination(f=2, chunk_size=512), "gradients"),
    "caf": (lambda: CAF(f=2, chunk_size=3), "gradients"),
    "clipping": (lambda: Clipping(threshold=40.0, chunk_size=3), "vectors"),
    "arc": (lambda: ARC(f=2, chunk_size=3), "vectors"),
    "nnm": (lambda: NearestNeighborMixing(f=2, feature_chunk_size=512), "vectors"),
    "bucketing": (lambda: Bucketing(bucket_size=3, feature_chunk_size=512, perm=[3, 1, 4, 0, 5, 9, 2, 6, 8, 7]), "vectors"),
    "empire": (lambda: EmpireAttack(scale=-1.5, chunk_size=3), "honest_grads"),
    "little": (lambda: LittleAttack(f=2, chunk_size=512), "honest_grads"),
    "gaussian": (lambda: GaussianAttack(mu=0.5, sigma=2.0, seed=11, chunk_size=512), "honest_grads"),
    "inf": (lambda: InfAttack(chunk_size=512), "honest_grads"),
    "mimic": (lambda: MimicAttack(epsilon=4, chunk_size=512), "honest_grads"),
    "sign-flip": (lambda: SignFlipAttack(scale=-2.0, chunk_size=512), "base_grad"),
}


def _direct(op, key, rows):
    if key == "gradients":
        return op.aggregate(rows)
    if key == "vectors":
        return op.pre_aggregate(rows)
    if key == "base_grad":
        return op.apply(base_grad=rows[0])
    return op.apply(honest_grads=rows)


def _same(a, b, tol):
    if isinstance(a, (list, tuple)):
        assert len(a) == len(b)
        return all(_same(x, y, tol) for x, y in zip(a, b))
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    return bool(torch.allclose(a, b, rtol=tol, atol=tol, equal_nan=True))


async def _through_pool(backend, names):
    rows = _data()
    pool = ActorPool([ActorPoolConfig(backend, count=2)])
    await pool.start()
    out = {}
    try:
        for name in names:
            mk, key = OPERATORS[name]
            graph = make_single_operator_graph(node_name="op", operator=mk(), input_keys=(key,))
            data = rows[0] if key == "base_grad" else rows
            out[name] = (await NodeScheduler(graph, pool=pool).run({key: data}))["op"]
    finally:
        await pool.shutdown()
    return out


@pytest.fixture
def reference_dispatch(monkeypatch):
    monkeypatch.setenv("BYZPY_POOL_DISPATCH", "reference")


def _check(results):
    rows = _data()
    bad = []
    for name, got in results.items():
        mk, key = OPERATORS[name]
        want = _direct(mk(), key, rows)
        # iterative solvers stop on a tolerance: the pool path accumulates the Gram matrix chunk by chunk
        tol = 2e-4 if name.startswith(("geometric-median", "caf", "centered", "smea")) else 2e-5
        if not _same(got, want, tol):
            bad.append(name)
    assert not bad, bad


def test_every_operator_on_thread_workers(reference_dispatch):
    _check(asyncio.run(_through_pool("thread", list(OPERATORS))))


def test_every_operator_on_process_workers_and_no_shared_memory_is_left_behind(reference_dispatch):
    before = set(glob.glob("/dev/shm/psm_*"))
    _check(asyncio.run(_through_pool("process", list(OPERATORS))))
    leaked = set(glob.glob("/dev/shm/psm_*")) - before
    assert not leaked, sorted(leaked)


def test_subtask_paths_were_really_taken(reference_dispatch, monkeypatch):
    """Guard against the test silently exercising the direct route: count ``create_subtasks`` /
    ``run_barriered_subtasks`` calls per operator class."""
    from byzpy_b200.engine.graph.operator import Operator

    calls = []
    orig_run = Operator._run_subtasks

    async def spy(self, pool, subtasks, limit, context):
        subtasks = list(subtasks)
        calls.append((type(self).__name__, len(subtasks)))
        return await orig_run(self, pool, subtasks, limit, context)

    monkeypatch.setattr(Operator, "_run_subtasks", spy)
    names = ["median", "multi-krum", "nnm", "bucketing", "empire", "little", "gaussian", "sign-flip"]
    asyncio.run(_through_pool("thread", names))
    seen = {cls for cls, k in calls if k >= 2}
    assert {"CoordinateWiseMedian", "MultiKrum", "NearestNeighborMixing", "Bucketing", "EmpireAttack", "LittleAttack",
            "GaussianAttack", "SignFlipAttack"} <= seen, calls
    assert max(k for _, k in calls) <= 16, calls          # a fan-out, not one subtask per handful of coordinates


def test_default_granularity_stays_sane_at_benchmark_sizes():
    """With the constructors' default ``chunk_size`` and the sizes of the reference's benchmark scripts (64 x 65 536, a
    4-worker pool) every operator creates between 1 and 64 subtasks.  (Empire's ``chunk_size`` counts gradients,
    reference attacks/empire.py:108: read as coordinates its default of 8 would mean 8192 subtasks.)"""
    from byzpy_b200.engine.graph.operator import OpContext

    n, d = 64, 65536
    rows = [torch.zeros(d) + i for i in range(n)]
    ctx = OpContext(node_name="x", metadata={"pool_size": 4, "pool_in_process": True})
    ops = {"median": (CoordinateWiseMedian(), "gradients"), "trimmed": (CoordinateWiseTrimmedMean(f=8), "gradients"),
           "meamed": (MeanOfMedians(f=8), "gradients"), "krum": (Krum(f=8), "gradients"),
           "multi-krum": (MultiKrum(f=8, q=8), "gradients"), "monna": (MoNNA(f=8), "gradients"),
           "cge": (ComparativeGradientElimination(f=8), "gradients"), "caf": (CAF(f=8), "gradients"),
           "clipping": (Clipping(), "vectors"), "arc": (ARC(f=8), "vectors"),
           "nnm": (NearestNeighborMixing(f=8), "vectors"), "bucketing": (Bucketing(bucket_size=4), "vectors"),
           "empire": (EmpireAttack(), "honest_grads"), "little": (LittleAttack(f=8), "honest_grads"),
           "gaussian": (GaussianAttack(seed=1), "honest_grads"), "inf": (InfAttack(), "honest_grads"),
           "mimic": (MimicAttack(), "honest_grads"), "sign-flip": (SignFlipAttack(), "base_grad")}
    counts = {}
    for name, (op, key) in ops.items():
        data = rows[0] if key == "base_grad" else rows
        counts[name] = len(list(op.create_subtasks({key: data}, context=ctx)))
    assert all(1 <= k <= 64 for k in counts.values()), counts


def test_random_bucketing_through_a_pool_is_a_valid_bucketing(reference_dispatch):
    rows = _data()

    async def run():
        pool = ActorPool([ActorPoolConfig("thread", count=2)])
        await pool.start()
        try:
            op = Bucketing(bucket_size=3, feature_chunk_size=512, rng=random.Random(7))
            graph = make_single_operator_graph(node_name="op", operator=op, input_keys=("vectors",))
            return (await NodeScheduler(graph, pool=pool).run({"vectors": rows}))["op"]
        finally:
            await pool.shutdown()

    out = asyncio.run(run())
    assert len(out) == 4
    total = sum(o * k for o, k in zip(out, (3, 3, 3, 1)))          # bucket means x bucket sizes = the sum of all rows
    assert torch.allclose(total, torch.stack(rows).sum(0), atol=1e-3)
