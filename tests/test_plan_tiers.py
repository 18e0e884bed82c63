"""Fused-round planning logic that is pure n-space arithmetic and therefore checkable without a
GPU: aggregator plans (weights over rows reproduce ``aggregate``), their composition with linear
pre-aggregators (``G' = W G W^T``), attack folds, row layouts, and the generic
parameter-server round's failure handling (reference engine/parameter_server/ps.py:28-262)."""
import asyncio

import numpy as np
import pytest
import torch

from byzpy_b200 import ops
from byzpy_b200.aggregators.coordinate_wise import (CoordinateWiseMedian, CoordinateWiseTrimmedMean,
                                                     MeanOfMedians)
from byzpy_b200.aggregators.geometric_wise import (SMEA, GeometricMedian, Krum,
                                                    MinimumDiameterAveraging, MoNNA, MultiKrum)
from byzpy_b200.aggregators.norm_wise import CAF, CenteredClipping, ComparativeGradientElimination
from byzpy_b200.attacks import (EmpireAttack, GaussianAttack, InfAttack, LittleAttack, MimicAttack,
                                SignFlipAttack)
from byzpy_b200.engine.parameter_server.ps import ParameterServer
from byzpy_b200.parallel.device_ps import CwPlan, GramPlan, MapCwPlan, RowFold, RowLayout
from byzpy_b200.pre_aggregators import ARC, Bucketing, Clipping, NearestNeighborMixing


def run(coro):
    return asyncio.run(coro)


def rows(n=9, d=37, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(d, generator=g) + (6.0 if i >= n - 2 else 0.0) for i in range(n)]


def apply_plan(plan, vs):
    """What the fused round does with a Gram plan, in plain torch."""
    X = torch.stack(vs).double()
    aux = [X.sort(0).values[(len(vs) - 1) // 2]] if "median" in plan.aux else []
    A = torch.cat([X, torch.stack(aux)]) if aux else X
    w = plan.solver(A @ A.T)
    assert w.dtype == torch.float32 and w.shape == (A.shape[0],)
    return (w.double() @ A).float()


GRAM = [lambda: MultiKrum(f=2, q=3), lambda: Krum(f=2), lambda: GeometricMedian(), lambda: GeometricMedian(init="mean"),
        lambda: MinimumDiameterAveraging(f=2), lambda: MoNNA(f=2, reference_index=1), lambda: SMEA(f=2),
        lambda: CenteredClipping(c_tau=0.7, M=5), lambda: CenteredClipping(c_tau=0.7, M=5, init="median"),
        lambda: CenteredClipping(c_tau=2.0, init="zero"), lambda: ComparativeGradientElimination(f=2)]


@pytest.mark.parametrize("mk", GRAM)
def test_gram_plan_weights_reproduce_aggregate(mk):
    vs = rows()
    plan = mk().fused_plan(len(vs))
    assert isinstance(plan, GramPlan) and plan.name == mk().name
    assert torch.allclose(apply_plan(plan, vs), mk().aggregate(vs), rtol=1e-4, atol=1e-4)


def test_gram_plan_aux_and_capturability_flags():
    assert GeometricMedian().fused_plan(5).aux == ("median",) and GeometricMedian(init="mean").fused_plan(5).aux == ()
    assert CenteredClipping(c_tau=1.0, init="median").fused_plan(5).aux == ("median",)
    assert MultiKrum(f=1, q=1).fused_plan(5).aux == ()
    for agg in (MultiKrum(f=1, q=2), GeometricMedian(), CenteredClipping(c_tau=1.0)):
        assert agg.fused_plan(6).capturable, agg.name
    # CAF's filter loop runs as a device n-space kernel; its plan carries one constant aux row (the reference's
    # fixed power-iteration start direction) and is graph-capturable
    caf = CAF(f=1).fused_plan(6)
    assert caf.capturable and len(caf.aux) == 1 and caf.aux[0][0] == "const"


@pytest.mark.parametrize("agg,mode,f", [(CoordinateWiseMedian(), ops.MODE_MEDIAN, 0),
                                        (CoordinateWiseTrimmedMean(f=2), ops.MODE_TRMEAN, 2),
                                        (MeanOfMedians(f=3), ops.MODE_MEAMED, 3)])
def test_coordinate_wise_plans(agg, mode, f):
    plan = agg.fused_plan(9)
    assert isinstance(plan, CwPlan) and plan.mode == mode and plan.f == f


@pytest.mark.parametrize("agg", [CoordinateWiseTrimmedMean(f=3), MeanOfMedians(f=6), MultiKrum(f=5, q=1), SMEA(f=3),
                                 MoNNA(f=3), MinimumDiameterAveraging(f=6), ComparativeGradientElimination(f=6)])
def test_plans_validate_the_row_count(agg):
    with pytest.raises(ValueError):
        agg.fused_plan(6)


# --------------------------------------------------------------- composition with pre-aggregators
class _Dev:
    """Minimal stand-in for a node: only ``device`` is consulted when composing plans."""

    device = torch.device("cpu")


def _ps(agg, pre, mapcw=True):
    ps = ParameterServer([_Dev()], [], agg, pre_aggregator=pre, fused=None)
    ps._allow_mapcw = mapcw          # what fused=True / BYZPY_FUSED_MAPCW=1 grant on a GPU box
    return ps


PRE = [lambda: Clipping(threshold=3.0), lambda: ARC(f=2), lambda: NearestNeighborMixing(f=2),
       lambda: Bucketing(bucket_size=2, perm=[3, 1, 4, 0, 5, 7, 2, 6, 8]), lambda: Bucketing(bucket_size=3, perm=range(9))]
INNER = [lambda: MultiKrum(f=1, q=2), lambda: GeometricMedian(init="mean"), lambda: CenteredClipping(c_tau=0.7, M=4),
         lambda: ComparativeGradientElimination(f=1), lambda: MinimumDiameterAveraging(f=1)]


@pytest.mark.parametrize("mk_pre", PRE)
@pytest.mark.parametrize("mk_agg", INNER)
def test_preaggregator_composes_in_n_space(mk_pre, mk_agg):
    vs = rows(9, 29, seed=3)
    if isinstance(mk_agg(), ComparativeGradientElimination) and isinstance(mk_pre(), (Clipping, ARC)):
        pytest.skip("clipped rows tie at exactly the threshold norm: which of them CGE drops is rounding noise")
    plan = _ps(mk_agg(), mk_pre())._fused_plan(len(vs))
    assert isinstance(plan, GramPlan) and plan.name == f"{mk_pre().name}+{mk_agg().name}" and plan.aux == ()
    if plan.refresh is not None:
        plan.refresh()
    expect = mk_agg().aggregate(mk_pre().pre_aggregate(vs))
    assert torch.allclose(apply_plan(plan, vs), expect, rtol=1e-4, atol=1e-4)
    # Clipping / ARC / NNM maps are device n-space kernels (csrc/nspace_maps.cu), Bucketing's map is a constant
    # matrix refreshed on the host before the round: the composed plan is capturable whenever the inner one is
    assert plan.capturable == mk_agg().fused_plan(3).capturable


CW_INNER = [lambda: CoordinateWiseMedian(), lambda: CoordinateWiseTrimmedMean(f=1), lambda: MeanOfMedians(f=1)]


@pytest.mark.parametrize("mk_pre", PRE)
@pytest.mark.parametrize("mk_agg", CW_INNER)
def test_preaggregator_with_a_coordinate_wise_aggregator_gives_a_map_plan(mk_pre, mk_agg):
    """Bucketing -> median, NNM -> trimmed mean, ...: not linear in n-space, so the plan carries the (m, n) map
    and the coordinate-wise plan of the m mixed rows; emulated here in plain torch."""
    vs = rows(9, 29, seed=5)
    pre, agg = mk_pre(), mk_agg()
    plan = _ps(agg, pre)._fused_plan(len(vs))
    assert isinstance(plan, MapCwPlan) and isinstance(plan.cw, CwPlan) and plan.name == f"{pre.name}+{agg.name}"
    assert plan.needs_gram == pre.needs_gram and plan.capturable
    if plan.refresh is not None:
        plan.refresh()
    X = torch.stack(vs).double()
    W = plan.weights((X @ X.T) if plan.needs_gram else None)
    mixed = mk_pre().pre_aggregate(vs)
    assert W.shape == (plan.m, len(vs)) and plan.m == len(mixed)
    Y = (W.double() @ X).float()
    torch.testing.assert_close(Y, torch.stack(list(mixed)), rtol=1e-5, atol=1e-5)
    expect = mk_agg().aggregate(list(mixed))
    got = ops.cw_select(list(Y.unbind(0)), plan.cw.mode, plan.cw.f)
    torch.testing.assert_close(got, expect, rtol=1e-5, atol=1e-5)
    assert plan.cw.mode == mk_agg().fused_plan(plan.m).mode and plan.cw.f == mk_agg().fused_plan(plan.m).f


def test_bucketing_plan_refresh_draws_a_new_permutation_each_round():
    import random

    vs = rows(8, 11, seed=4)
    pre = Bucketing(bucket_size=2, rng=random.Random(0))
    plan = _ps(MultiKrum(f=0, q=4), pre)._fused_plan(8)             # q = all buckets: result is the grand mean
    outs = []
    for _ in range(3):
        plan.refresh()
        outs.append(apply_plan(plan, vs))
    mean = torch.stack(vs).mean(0)
    assert all(torch.allclose(o, mean, atol=1e-4) for o in outs)
    Ws = []
    for _ in range(4):
        plan.refresh()
        X = torch.eye(8, dtype=torch.float64)
        Ws.append(plan.solver(X @ X.T).clone())
    assert all(abs(float(w.sum()) - 1.0) < 1e-5 for w in Ws)


def test_composition_is_refused_when_it_cannot_be_expressed():
    from byzpy_b200.pre_aggregators.base import PreAggregator

    class Opaque(PreAggregator):
        name = "opaque"

        def pre_aggregate(self, xs):
            return list(xs)

    assert _ps(MultiKrum(f=1, q=1), Opaque())._fused_plan(6) is None            # not a linear map
    assert isinstance(_ps(CoordinateWiseMedian(), Clipping())._fused_plan(6), MapCwPlan)   # mixed rows per shard
    # ... only on request: automatic selection keeps to the plans that have been timed on hardware
    assert _ps(CoordinateWiseMedian(), Clipping(), mapcw=False)._fused_plan(6) is None
    assert _ps(CoordinateWiseMedian(), Opaque())._fused_plan(6) is None
    assert _ps(GeometricMedian(), Clipping())._fused_plan(6) is None             # median start row is not linear in W
    with pytest.raises(ValueError):
        _ps(MultiKrum(f=1, q=1), NearestNeighborMixing(f=6))._fused_plan(6)
    assert isinstance(_ps(MultiKrum(f=1, q=1), None)._fused_plan(6), GramPlan)


# -------------------------------------------------------------------------------------- attack folds
def test_attack_folds():
    assert SignFlipAttack(scale=-4.0).fold(5) == RowFold("scale", scale=-4.0)
    e = EmpireAttack(scale=-1.5).fold(7)
    assert e.kind == "virtual" and e.a == -1.5 and e.b == 0.0
    l = LittleAttack(f=2, N=9).fold(7)
    assert l.kind == "virtual" and l.a == 1.0 and l.b != 0.0
    m = MimicAttack(epsilon=2).fold(5)
    assert m.kind == "alias" and m.index == 2
    assert GaussianAttack().fold(5) is None and InfAttack().fold(5) is None


def test_virtual_fold_matches_the_materialised_attack():
    vs = rows(7, 13, seed=5)
    X = torch.stack(vs).double()
    mu, sd = X.mean(0), X.std(0, unbiased=False)
    for atk in (EmpireAttack(scale=-2.0), LittleAttack(f=2, N=9)):
        f = atk.fold(len(vs))
        assert torch.allclose((f.a * mu + f.b * sd).float(), atk.apply(honest_grads=vs), rtol=1e-4, atol=1e-4), atk.name


# ---------------------------------------------------------------------------------------- row layouts
def test_block_layout():
    lay = RowLayout.block(6, 2, 4, n_virtual=1)
    assert lay.n_workers == 8 and lay.rank_of == [0, 0, 1, 1, 2, 2, 3, 3] and lay.slot_of == [0, 1] * 4
    assert lay.local_ids(2) == [4, 5] and lay.max_local() == 2 and lay.n_virtual == 1
    with pytest.raises(ValueError, match="divide evenly"):
        RowLayout.block(5, 0, 2)
    one = RowLayout.block(3, 1, 1)
    assert one.local_ids(0) == [0, 1, 2, 3] and one.max_local() == 4


@pytest.mark.parametrize("n,world", [(10, 4), (3, 8), (8, 8), (1, 2), (0, 2)])
def test_spread_layout_is_balanced_and_contiguous(n, world):
    lay = RowLayout.spread(n, 0, world)
    sizes = [len(lay.local_ids(r)) for r in range(world)]
    assert sum(sizes) == n and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert lay.rank_of == sorted(lay.rank_of) and lay.max_local() == max(sizes)
    assert all(lay.slot_of[g] == lay.local_ids(lay.rank_of[g]).index(g) for g in range(n))


# -------------------------------------------------------------------- generic round: failure handling
class _Node:
    def __init__(self, grad, delay=0.0, fail=False):
        self.grad, self.delay, self.fail = torch.tensor(grad), delay, fail
        self.applied = []

    async def honest_gradient_for_next_batch(self):
        await asyncio.sleep(self.delay)
        if self.fail:
            raise RuntimeError("node crashed")
        return self.grad

    def byzantine_gradient_for_next_batch(self, honest):
        if self.fail:
            raise RuntimeError("byz crashed")
        return -torch.stack(list(honest)).mean(0)

    def apply_server_gradient(self, g):
        self.applied.append(g.clone())


def test_round_collects_sync_and_async_nodes_and_applies_to_honest_only():
    hon, byz = [_Node([1.0, 1.0]), _Node([3.0, 3.0])], [_Node([0.0, 0.0])]
    ps = ParameterServer(hon, byz, CoordinateWiseMedian())
    assert ps.device_round is None
    g = ps.round_sync()
    assert torch.equal(g, torch.tensor([1.0, 1.0])) and ps.rounds == 1     # rows 1, 3, -2 -> lower median 1
    assert all(len(n.applied) == 1 for n in hon) and byz[0].applied == []
    ps2 = ParameterServer(hon, byz, CoordinateWiseMedian(), update_byzantines=True)
    ps2.round_sync()
    assert len(byz[0].applied) == 1
    with pytest.raises(RuntimeError, match="fused device path"):
        ps.step()


def test_failures_propagate_unless_tolerated():
    hon = [_Node([1.0]), _Node([2.0], fail=True), _Node([3.0])]
    with pytest.raises(RuntimeError, match="node crashed"):
        ParameterServer(hon, [], CoordinateWiseMedian()).round_sync()
    ps = ParameterServer(hon, [_Node([0.0], fail=True)], CoordinateWiseMedian(), tolerate_failures=True)
    g = ps.round_sync()
    assert torch.equal(g, torch.tensor([1.0]))                              # survivors 1, 3 -> lower median
    assert [(r, who) for r, who, _ in ps.failed] == [(0, "honest:1"), (0, "byzantine:0")]
    assert "node crashed" in ps.failed[0][2]


def test_slow_nodes_time_out_as_silent():
    hon = [_Node([1.0]), _Node([9.0], delay=1.0), _Node([2.0])]
    ps = ParameterServer(hon, [], CoordinateWiseMedian(), node_timeout=0.05, tolerate_failures=True)
    assert torch.equal(ps.round_sync(), torch.tensor([1.0]))
    assert ps.failed[0][1] == "honest:1" and "Timeout" in ps.failed[0][2]
    with pytest.raises(asyncio.TimeoutError):
        ParameterServer(hon, [], CoordinateWiseMedian(), node_timeout=0.05).round_sync()


def test_all_nodes_failed_is_an_error():
    ps = ParameterServer([_Node([1.0], fail=True)], [], CoordinateWiseMedian(), tolerate_failures=True)
    with pytest.raises(RuntimeError, match="all nodes failed"):
        ps.round_sync()


def test_fused_true_without_device_nodes_is_refused():
    with pytest.raises(RuntimeError, match="no fused device path"):
        ParameterServer([_Node([1.0])], [], CoordinateWiseMedian(), fused=True)


def test_shutdown_closes_actor_backed_nodes():
    closed = []

    class Backend:
        async def close(self):
            closed.append(1)

    class Ref:
        _backend = Backend()

    n = _Node([1.0])
    n._ref = Ref()
    run(ParameterServer([n, _Node([2.0])], [], CoordinateWiseMedian()).shutdown())
    assert closed == [1]


# ------------------------------------------------------------------- sync-free selection solvers
@pytest.mark.parametrize("seed", range(4))
def test_torch_selection_solvers_match_the_host_solvers(seed):
    from byzpy_b200.ops import nspace, nspace_cuda

    X = torch.stack(rows(11, 17, seed=seed)).double()
    X[3] = X[7]                                                     # an exact tie
    A = torch.cat([X, X.mean(0, keepdim=True)])                     # one aux row the solvers must ignore
    G = A @ A.T
    n = X.shape[0]
    for f in (0, 2, 5):
        w = nspace_cuda.cge_weights(G, n, f)
        assert w.dtype == torch.float32 and w.shape == (n + 1,) and w[n] == 0
        assert np.allclose(w[:n].numpy(), nspace.cge_weights(G[:n, :n].numpy(), f))
        for ref in (0, 3, 10):
            wm = nspace_cuda.monna_weights(G, n, f, ref)
            assert wm[n] == 0 and np.allclose(wm[:n].numpy(), nspace.monna_weights(G[:n, :n].numpy(), f, ref))


def test_torch_selection_solvers_treat_nan_rows_as_farthest():
    from byzpy_b200.ops import nspace, nspace_cuda

    X = torch.stack(rows(6, 5, seed=9)).double()
    G = X @ X.T
    G[2, :] = float("nan")
    G[:, 2] = float("nan")
    assert nspace_cuda.cge_weights(G, 6, 1)[2] == 0 and nspace_cuda.monna_weights(G, 6, 1, 0)[2] == 0
    assert np.allclose(nspace_cuda.cge_weights(G, 6, 1).numpy(), nspace.cge_weights(G.numpy(), 1))
    assert np.allclose(nspace_cuda.monna_weights(G, 6, 1, 0).numpy(), nspace.monna_weights(G.numpy(), 1, 0))


def test_cge_and_monna_plans_are_capturable():
    assert ComparativeGradientElimination(f=1).fused_plan(6).capturable and MoNNA(f=1).fused_plan(6).capturable


# ------------------------------------------------------------------ gradient buckets (CPU)
def test_bucket_planner_cuts_resnet18_at_block_inputs_from_the_tail():
    import torch

    from byzpy_b200.models import resnet18
    from byzpy_b200.ops.fused_layers import bucket_candidates
    from byzpy_b200.parallel.arena import ParamArena
    from byzpy_b200.parallel.device_ps import bucket_bounds, pick_bucket_offsets

    m = resnet18(num_classes=1000)
    a = ParamArena(m)
    cand = bucket_candidates(m, a.offsets)
    offs = [fo for fo, _ in cand]
    assert offs == sorted(offs) and len(offs) == 11          # stem conv, bn1, 8 blocks, fc
    picks = pick_bucket_offsets(offs, a.d, (0.40, 0.75, 0.93))
    names = {id(mod): n for n, mod in m.named_modules()}
    assert [names[id(mod)] for fo, mod in cand if fo in picks] == ["layer3.0", "layer4.0", "layer4.1"]
    b = bucket_bounds(picks, a.d_pad, 1 << 16)
    assert b[0] == a.d_pad and b[-1] == 0 and b == sorted(b, reverse=True)
    assert all(x % 1024 == 0 for x in b)
    assert all(bound >= off for bound, off in zip(b[1:-1], picks))      # rounded UP: never early
    assert pick_bucket_offsets(offs, a.d, (0.40, 0.75, 0.93), buckets=2) == picks[:1]
    assert pick_bucket_offsets(offs, a.d, (0.40, 0.75, 0.93), buckets=1) == []
    # tiny buckets are merged away
    assert bucket_bounds([5000, 3000, 100], 8192, 4096) == [8192, 0]
    assert torch.is_tensor(a.flat_grads)


def test_bucket_marks_fire_in_reverse_layer_order_with_complete_gradients():
    import torch

    from byzpy_b200.models import resnet18
    from byzpy_b200.ops.fused_layers import bucket_candidates, install_bucket_marks
    from byzpy_b200.parallel.arena import ParamArena
    from byzpy_b200.parallel.device_ps import pick_bucket_offsets

    torch.manual_seed(0)
    m = resnet18(num_classes=10, small_input=True)
    a = ParamArena(m)
    cand = bucket_candidates(m, a.offsets)
    picks = pick_bucket_offsets([fo for fo, _ in cand], a.d, (0.40, 0.75, 0.93))
    seen = []

    def on_mark(k):
        g = a.flat_grads
        done = float((g[picks[k]:a.d] != 0).float().mean())
        early = float((g[:picks[k]] != 0).float().mean())
        seen.append((k, done, early))

    handles = install_bucket_marks([(next(mod for off, mod in cand if off == fo), k) for k, fo in enumerate(picks)],
                                   on_mark)
    a.zero_grad()
    m(torch.randn(4, 3, 32, 32)).sum().backward()
    assert [k for k, _, _ in seen] == [0, 1, 2]
    for _, done, early in seen:
        assert done > 0.99 and early == 0.0       # everything behind the mark exists, nothing before it
    for h in handles:
        h.remove()
    seen.clear()
    a.zero_grad()
    m(torch.randn(4, 3, 32, 32)).sum().backward()
    assert seen == []


def test_coordinate_shards_follow_the_live_set_and_equal_the_static_split_when_nobody_was_dropped():
    from byzpy_b200.parallel.device_ps import DeviceRound

    def shard(rank, world, d_pad, live_mask=0):
        r = object.__new__(DeviceRound)
        r.rank, r.world, r.live_mask = rank, world, live_mask
        return r._shard(0, d_pad)

    for world in (1, 2, 3, 4, 5, 8):
        for d_pad in (1024, 11_689_984, 25_558_016):
            sh = d_pad // world
            sh -= sh % 4
            cover = 0
            for rank in range(world):
                static = (rank * sh, sh if rank < world - 1 else d_pad - sh * (world - 1))     # DeviceRound.__init__
                assert shard(rank, world, d_pad) == static
                assert static[0] % 4 == 0 and static[1] % 4 == 0
                cover += static[1]
            assert cover == d_pad
    # after rank 1 of 4 was dropped the survivors split the whole vector among themselves
    live = 0b1101
    parts = [shard(r, 4, 4096, live) for r in (0, 2, 3)]
    assert parts[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(parts, parts[1:])) and sum(p[1] for p in parts) == 4096
