"""Fine-grained unit tiers of the scheduling layer, one behaviour per test: graph structure, lazy
builder, execution sessions / futures, the one-operator executor, the dataflow scheduler, the
message-aware scheduler, subtasks and the windowed dispatcher (mirrors the reference tiers
tests/engine/graph/test_{graph,lazy,session,executor,parallel_scheduler,scheduler}.py)."""
import asyncio
import time

import pytest
import torch

from byzpy_b200 import OperatorExecutor, run_operator
from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean
from byzpy_b200.aggregators.geometric_wise import MultiKrum
from byzpy_b200.attacks import EmpireAttack, SignFlipAttack
from byzpy_b200.engine.graph.graph import ComputationGraph, GraphInput, GraphNode, graph_input
from byzpy_b200.engine.graph.lazy import GraphBuilder, LazyNode
from byzpy_b200.engine.graph.operator import (MessageTriggerOp, OpContext, Operator, _window_size,
                                               run_subtasks_windowed)
from byzpy_b200.engine.graph.ops import CallableOp, RemoteCallableOp, make_single_operator_graph
from byzpy_b200.engine.graph.parallel_scheduler import ParallelScheduler
from byzpy_b200.engine.graph.scheduler import MessageAwareNodeScheduler, MessageSource, NodeScheduler
from byzpy_b200.engine.graph.session import ExecutionFuture, ExecutionSession
from byzpy_b200.engine.graph.subtask import SubTask
from byzpy_b200.pre_aggregators import Bucketing, Clipping


def run(coro):
    return asyncio.run(coro)


class Plus(Operator):
    """Sums its inputs and adds ``k``; counts invocations."""

    name = "plus"

    def __init__(self, k=0):
        self.k, self.calls = k, 0

    def compute(self, inputs, *, context):
        self.calls += 1
        return sum(inputs.values()) + self.k


class Nap(Operator):
    name = "nap"

    def __init__(self, log, dt=0.03):
        self.log, self.dt = log, dt

    async def compute(self, inputs, *, context):
        self.log.append(("start", context.node_name, time.perf_counter()))
        await asyncio.sleep(self.dt)
        self.log.append(("end", context.node_name, time.perf_counter()))
        return context.node_name


class Boom(Operator):
    name = "boom"

    def compute(self, inputs, *, context):
        raise ZeroDivisionError("boom")


def _node(name, op=None, **inputs):
    return GraphNode(name, op or Plus(), inputs)


def _times(log, kind, name):
    return [t for k, n, t in log if k == kind and n == name][0]


# ------------------------------------------------------------------------------- graph structure
def test_graph_requires_nodes():
    with pytest.raises(ValueError, match="at least one node"):
        ComputationGraph([])


def test_graph_rejects_duplicate_names():
    with pytest.raises(ValueError, match="Duplicate"):
        ComputationGraph([_node("a", x=graph_input("x")), _node("a", x=graph_input("x"))])


def test_graph_rejects_unknown_dependency():
    with pytest.raises(ValueError, match="unknown node"):
        ComputationGraph([_node("a", x="ghost")])


def test_graph_rejects_unknown_output():
    with pytest.raises(ValueError, match="Unknown output"):
        ComputationGraph([_node("a", x=graph_input("x"))], outputs=["b"])


def test_graph_rejects_two_cycle_and_self_loop():
    with pytest.raises(ValueError, match="cycle"):
        ComputationGraph([_node("a", x="b"), _node("b", x="a")])
    with pytest.raises(ValueError, match="cycle"):
        ComputationGraph([_node("a", x="a")])


def test_graph_default_output_is_last_in_topological_order():
    g = ComputationGraph([_node("z", x="a"), _node("a", x=graph_input("x"))])
    assert [n.name for n in g.nodes_in_order()] == ["a", "z"]
    assert g.outputs == ["z"]


def test_graph_order_is_stable_for_independent_nodes():
    g = ComputationGraph([_node(n, x=graph_input("x")) for n in "dcba"], outputs=["a"])
    assert [n.name for n in g.nodes_in_order()] == list("dcba")


def test_graph_required_inputs_and_dependencies():
    g = ComputationGraph([_node("a", x=graph_input("p"), y=graph_input("q")),
                          _node("b", x="a", y="a", z=graph_input("p"))])
    assert g.required_inputs == frozenset({"p", "q"})
    assert g.dependencies("b") == ["a"] and g.dependencies("a") == []
    assert len(g) == 2 and "a" in g and "zz" not in g and g.node("b").name == "b"


def test_graph_message_source_is_neither_input_nor_edge():
    src = GraphInput.from_message("grad", field="v", timeout=1.5)
    assert isinstance(src, MessageSource) and (src.message_type, src.field, src.timeout) == ("grad", "v", 1.5)
    g = ComputationGraph([_node("a", x=src, y=graph_input("y"))])
    assert g.required_inputs == frozenset({"y"}) and g.dependencies("a") == []
    assert "grad" in repr(src)


def test_graph_input_is_hashable_value_object():
    assert graph_input("x") == GraphInput("x") and len({graph_input("x"), GraphInput("x")}) == 1
    with pytest.raises(Exception):
        graph_input("x").name = "y"


def test_diamond_graph_values():
    g = ComputationGraph([_node("src", Plus(1), x=graph_input("x")), _node("l", Plus(10), x="src"),
                          _node("r", Plus(100), x="src"), _node("sink", Plus(), a="l", b="r")])
    for sched in (NodeScheduler, ParallelScheduler):
        assert run(sched(g).run({"x": 1})) == {"sink": (2 + 10) + (2 + 100)}


def test_multiple_outputs_returned_in_declared_order():
    g = ComputationGraph([_node("a", Plus(1), x=graph_input("x")), _node("b", Plus(2), x="a")], outputs=["b", "a"])
    out = run(NodeScheduler(g).run({"x": 0}))
    assert list(out) == ["b", "a"] and out == {"b": 3, "a": 1}


# ------------------------------------------------------------------------------------ lazy builder
def test_lazy_input_is_idempotent_and_flagged():
    b = GraphBuilder()
    x = b.input("vectors")
    assert isinstance(x, LazyNode) and x.key == "vectors" and x._is_input
    assert b.input("vectors").key == x.key


def test_lazy_generated_names_are_unique_with_running_counter():
    b = GraphBuilder()
    x = b.input("vectors")
    a = x.apply(Clipping(threshold=1.0))
    c = a.apply(Clipping(threshold=2.0))
    m = c.apply(CoordinateWiseMedian())
    assert a.key.endswith("_0") and c.key.endswith("_1") and a.key[:-2] == c.key[:-2]
    assert m.key.endswith("_2") and m.key.startswith(CoordinateWiseMedian.name) and len({a.key, c.key, m.key}) == 3


def test_lazy_explicit_name_and_default_output():
    b = GraphBuilder()
    out = b.input("vectors").apply(CoordinateWiseMedian(), name="agg")
    g = b.build(outputs=["agg"])
    assert out.key == "agg" and g.outputs == ["agg"]
    with pytest.raises(TypeError):
        b.build()                      # outputs are mandatory, as in the reference builder


def test_lazy_input_key_defaults_to_operator_input_key():
    b = GraphBuilder()
    node = b.input("grads").apply(CoordinateWiseMedian(), name="m")
    wiring = b._nodes["m"].inputs
    assert list(wiring) == [CoordinateWiseMedian.input_key] and wiring[CoordinateWiseMedian.input_key] == GraphInput("grads")
    assert node.key == "m"


def test_lazy_custom_input_key_and_extra_inputs():
    b = GraphBuilder()
    x, y = b.input("x"), b.input("y")
    s = x.apply(Plus(), input_key="a", extra_inputs={"b": y}, name="s")
    t = s.apply(Plus(5), input_key="a", extra_inputs={"b": x, "c": s}, name="t")
    g = b.build(outputs=[t.key])
    assert g.required_inputs == frozenset({"x", "y"})
    assert run(NodeScheduler(g).run({"x": 1, "y": 2})) == {"t": 3 + 1 + 3 + 5}


def test_lazy_pipeline_matches_eager_composition():
    torch.manual_seed(0)
    vs = [torch.randn(33) * 4 for _ in range(8)]
    perm = [3, 1, 4, 0, 5, 7, 2, 6]
    b = GraphBuilder()
    out = b.input("vectors").apply(Clipping(threshold=2.0)).apply(Bucketing(bucket_size=2, perm=perm)) \
        .apply(CoordinateWiseTrimmedMean(f=1), name="final")
    g = b.build(outputs=[out.key])
    assert len(g) == 3
    torch.manual_seed(1)
    got = run(NodeScheduler(g).run({"vectors": vs}))["final"]
    torch.manual_seed(1)
    exp = CoordinateWiseTrimmedMean(f=1).aggregate(
        Bucketing(bucket_size=2, perm=perm).pre_aggregate(Clipping(threshold=2.0).pre_aggregate(vs)))
    assert torch.allclose(got, exp)


def test_lazy_fan_out_shares_one_upstream_node():
    b = GraphBuilder()
    clip = Plus(1)
    mid = b.input("x").apply(clip, input_key="v", name="mid")
    l = mid.apply(Plus(10), input_key="v", name="l")
    r = mid.apply(Plus(20), input_key="v", name="r")
    g = b.build(outputs=[l.key, r.key])
    assert run(ParallelScheduler(g).run({"x": 0})) == {"l": 11, "r": 21} and clip.calls == 1


def test_lazy_errors():
    b = GraphBuilder()
    x = b.input("x")
    with pytest.raises(TypeError, match="Operator"):
        x.apply(lambda v: v)
    with pytest.raises(ValueError):
        GraphBuilder().build(outputs=[])
    x.apply(Plus(), input_key="v", name="n")
    with pytest.raises(ValueError):
        b.build(outputs=["nope"])


def test_lazy_builder_is_reusable_after_build():
    b = GraphBuilder()
    n1 = b.input("x").apply(Plus(1), input_key="v", name="n1")
    g1 = b.build(outputs=[n1.key])
    n2 = n1.apply(Plus(1), input_key="v", name="n2")
    g2 = b.build(outputs=[n2.key])
    assert len(g1) == 1 and len(g2) == 2
    assert run(NodeScheduler(g2).run({"x": 0})) == {"n2": 2}


# -------------------------------------------------------------------------------- sessions/futures
def _chain(ops=None):
    a, b = ops or (Plus(1), Plus(1))
    return ComputationGraph([_node("a", a, x=graph_input("x")), _node("b", b, x="a")]), a, b


def test_session_caches_every_computed_node():
    g, a, b = _chain()

    async def go():
        s = ExecutionSession()
        assert await s.execute(g, {"x": 0}) == {"b": 2}
        assert s.is_cached("a") and s.is_cached("b") and s.get_cached("a") == 1
        assert not s.is_cached("x")      # inputs are not cached

    run(go())


def test_session_second_run_is_served_from_cache():
    g, a, b = _chain()

    async def go():
        s = ExecutionSession()
        await s.execute(g, {"x": 0})
        assert await s.execute(g, {"x": 50}) == {"b": 2}
        assert (a.calls, b.calls) == (1, 1)

    run(go())


def test_session_partial_reuse_across_graphs():
    shared = Plus(1)
    g1 = ComputationGraph([_node("a", shared, x=graph_input("x")), _node("b", Plus(1), x="a")])
    tail = Plus(10)
    g2 = ComputationGraph([_node("a", shared, x=graph_input("x")), _node("c", tail, x="a", y=graph_input("y"))])

    async def go():
        s = ExecutionSession()
        await s.execute(g1, {"x": 0})
        assert await s.execute(g2, {"x": 999, "y": 5}) == {"c": 1 + 5 + 10}
        assert shared.calls == 1 and tail.calls == 1

    run(go())


def test_session_cached_output_with_uncached_sibling():
    g = ComputationGraph([_node("a", Plus(1), x=graph_input("x")), _node("b", Plus(2), x=graph_input("x"))],
                         outputs=["a", "b"])
    only_a = ComputationGraph([_node("a", Plus(1), x=graph_input("x"))])

    async def go():
        s = ExecutionSession()
        await s.execute(only_a, {"x": 0})
        assert await s.execute(g, {"x": 0}) == {"a": 1, "b": 2}

    run(go())


def test_session_clear_cache_forces_recompute():
    g, a, b = _chain()

    async def go():
        s = ExecutionSession()
        await s.execute(g, {"x": 0})
        s.clear_cache()
        assert await s.execute(g, {"x": 5}) == {"b": 7}
        assert (a.calls, b.calls) == (2, 2)

    run(go())


def test_session_get_cached_missing_raises_keyerror():
    with pytest.raises(KeyError, match="No cached result"):
        ExecutionSession().get_cached("nothing")


def test_session_context_manager_clears_on_exit_even_on_error():
    g, _, _ = _chain()

    async def go():
        with pytest.raises(RuntimeError):
            async with ExecutionSession() as s:
                await s.execute(g, {"x": 0})
                assert s.is_cached("b")
                raise RuntimeError("user error")
        assert not s._result_cache

    run(go())


def test_session_without_cache_recomputes():
    g, a, b = _chain()

    async def go():
        s = ExecutionSession(cache_intermediate=False)
        assert await s.execute(g, {"x": 0}) == {"b": 2}
        assert await s.execute(g, {"x": 1}) == {"b": 3}
        assert a.calls == 2 and not s._result_cache

    run(go())


def test_session_metadata_reaches_operators():
    seen = {}

    class Peek(Operator):
        name = "peek"

        def compute(self, inputs, *, context):
            seen.update(context.metadata)
            return 0

    g = ComputationGraph([_node("p", Peek(), x=graph_input("x"))])
    run(ExecutionSession(metadata={"round": 7}).execute(g, {"x": 0}))
    assert seen["round"] == 7


def test_session_error_leaves_cache_unchanged():
    g = ComputationGraph([_node("a", Plus(1), x=graph_input("x")), _node("b", Boom(), x="a")])

    async def go():
        s = ExecutionSession()
        with pytest.raises(ZeroDivisionError):
            await s.execute(g, {"x": 0})
        assert not s.is_cached("b")

    run(go())


def test_session_missing_input_raises():
    g, _, _ = _chain()
    with pytest.raises(ValueError, match="Missing graph inputs"):
        run(ExecutionSession().execute(g, {}))


def test_future_await_wait_done_and_keys():
    g, _, _ = _chain()

    async def go():
        s = ExecutionSession()
        fut = s.execute_async(g, {"x": 1})
        assert isinstance(fut, ExecutionFuture) and fut.output_keys == ("b",) and not fut.done()
        assert await fut.wait() == {"b": 3}
        assert fut.done() and not fut.cancelled() and (await fut) == {"b": 3}

    run(go())


def test_future_cancel():
    log = []
    g = ComputationGraph([_node("slow", Nap(log, 5.0), x=graph_input("x"))])

    async def go():
        fut = ExecutionSession().execute_async(g, {"x": 0})
        await asyncio.sleep(0.01)
        assert fut.cancel()
        with pytest.raises(asyncio.CancelledError):
            await fut
        assert fut.cancelled()

    run(go())


def test_future_result_refuses_inside_running_loop_and_works_outside():
    g, _, _ = _chain()
    loop = asyncio.new_event_loop()
    try:
        async def make():
            return ExecutionSession().execute_async(g, {"x": 0})

        fut = loop.run_until_complete(make())

        async def inside():
            with pytest.raises(RuntimeError, match="running event loop"):
                fut.result()

        loop.run_until_complete(inside())
        assert fut.result(timeout=5) == {"b": 2}      # drives the owning loop to completion
        assert fut.result() == {"b": 2}               # already done
    finally:
        loop.close()


def test_futures_run_concurrently():
    log = []
    g1 = ComputationGraph([_node("n1", Nap(log, 0.05), x=graph_input("x"))])
    g2 = ComputationGraph([_node("n2", Nap(log, 0.05), x=graph_input("x"))])

    async def go():
        s = ExecutionSession(cache_intermediate=False)
        f1, f2 = s.execute_async(g1, {"x": 0}), s.execute_async(g2, {"x": 0})
        await asyncio.gather(f1.wait(), f2.wait())

    run(go())
    assert _times(log, "start", "n2") < _times(log, "end", "n1")


def test_future_propagates_exception():
    g = ComputationGraph([_node("b", Boom(), x=graph_input("x"))])

    async def go():
        fut = ExecutionSession().execute_async(g, {"x": 0})
        with pytest.raises(ZeroDivisionError):
            await fut

    run(go())


# ---------------------------------------------------------------------------------- one-op executor
def _vs(n=6, d=17, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(d, generator=g) for _ in range(n)]


def test_executor_autodetects_aggregator_key():
    ex = OperatorExecutor(CoordinateWiseMedian())
    assert ex.input_keys == (CoordinateWiseMedian.input_key,) and ex.node_name == CoordinateWiseMedian.name
    assert torch.equal(run(ex.run({CoordinateWiseMedian.input_key: _vs()})), CoordinateWiseMedian().aggregate(_vs()))


def test_executor_autodetects_preaggregator_key():
    out = run(OperatorExecutor(Clipping(threshold=0.5)).run({Clipping.input_key: _vs()}))
    exp = Clipping(threshold=0.5).pre_aggregate(_vs())
    assert len(out) == len(exp) and all(torch.allclose(a, b) for a, b in zip(out, exp))


def test_executor_custom_key_is_renamed():
    ex = OperatorExecutor(MultiKrum(f=1, q=2), input_keys=("my_grads",))
    assert ex._needs_input_mapping
    assert torch.allclose(run(ex.run({"my_grads": _vs()})), MultiKrum(f=1, q=2).aggregate(_vs()))
    renamed = ex._make_scheduler().graph.node(ex.node_name).op
    with pytest.raises(KeyError, match="my_grads"):
        run(renamed.run({"other": 1}, context=OpContext("n"), pool=None))


def test_executor_missing_input_raises_valueerror():
    with pytest.raises(ValueError, match="Missing graph inputs"):
        run(OperatorExecutor(CoordinateWiseMedian()).run({}))


def test_executor_custom_node_name():
    ex = OperatorExecutor(CoordinateWiseMedian(), node_name="agg42")
    run(ex.run({CoordinateWiseMedian.input_key: _vs()}))
    assert ex.node_name == "agg42" and ex._graph.outputs == ["agg42"]


def test_executor_attack_needs_explicit_keys():
    with pytest.raises(ValueError, match="Attack"):
        OperatorExecutor(EmpireAttack())
    out = run(OperatorExecutor(SignFlipAttack(), input_keys=("base_grad",)).run({"base_grad": torch.ones(4)}))
    assert torch.equal(out, -torch.ones(4))


def test_executor_rejects_non_operator_and_unknown_operator_kind():
    with pytest.raises(TypeError):
        OperatorExecutor(lambda x: x)
    with pytest.raises(ValueError, match="input_keys"):
        OperatorExecutor(Plus())
    assert run(OperatorExecutor(Plus(1), input_keys=("a", "b")).run({"a": 1, "b": 2})) == 4


def test_executor_is_reusable_across_runs():
    ex = OperatorExecutor(CoordinateWiseMedian())
    for seed in range(3):
        vs = _vs(seed=seed)
        assert torch.equal(run(ex.run({CoordinateWiseMedian.input_key: vs})), CoordinateWiseMedian().aggregate(vs))


def test_executor_owns_pool_inside_context_manager():
    from byzpy_b200.engine.graph.pool import ActorPoolConfig

    async def go():
        async with OperatorExecutor(CoordinateWiseMedian(chunk_size=8),
                                    pool_config=ActorPoolConfig("thread", count=2)) as ex:
            assert ex._pool is not None and ex._pool.size == 2
            first = await ex.run({CoordinateWiseMedian.input_key: _vs(d=40)})
            sched = ex._scheduler
            await ex.run({CoordinateWiseMedian.input_key: _vs(d=40)})
            assert ex._scheduler is sched          # scheduler reused while the pool lives
        assert ex._pool is None and ex._scheduler is None
        return first

    assert torch.allclose(run(go()), CoordinateWiseMedian().aggregate(_vs(d=40)))


def test_run_operator_with_pool_config_list():
    from byzpy_b200.engine.graph.pool import ActorPoolConfig

    out = run(run_operator(CoordinateWiseTrimmedMean(f=1, chunk_size=16), {"gradients": _vs(d=50)},
                           pool_config=[ActorPoolConfig("thread", count=1), ActorPoolConfig("thread", count=1)]))
    assert torch.allclose(out, CoordinateWiseTrimmedMean(f=1).aggregate(_vs(d=50)))


# ----------------------------------------------------------------------------- dataflow scheduler
def test_parallel_chain_is_sequential():
    log = []
    g = ComputationGraph([_node("a", Nap(log), x=graph_input("x")), _node("b", Nap(log), x="a"),
                          _node("c", Nap(log), x="b")])
    assert run(ParallelScheduler(g).run({"x": 0})) == {"c": "c"}
    assert _times(log, "end", "a") <= _times(log, "start", "b") <= _times(log, "end", "b") <= _times(log, "start", "c")


def test_parallel_independent_nodes_overlap():
    log = []
    g = ComputationGraph([_node(n, Nap(log, 0.1), x=graph_input("x")) for n in "abcd"], outputs=list("abcd"))
    t0 = time.perf_counter()
    run(ParallelScheduler(g).run({"x": 0}))
    assert time.perf_counter() - t0 < 0.35          # four 0.1 s naps back to back would take 0.4 s
    assert max(_times(log, "start", n) for n in "abcd") < min(_times(log, "end", n) for n in "abcd")


def test_parallel_max_concurrent_nodes_one_serialises():
    log = []
    g = ComputationGraph([_node(n, Nap(log, 0.02), x=graph_input("x")) for n in "abc"], outputs=list("abc"))
    run(ParallelScheduler(g, max_concurrent_nodes=1).run({"x": 0}))
    spans = sorted((_times(log, "start", n), _times(log, "end", n)) for n in "abc")
    assert all(spans[i][1] <= spans[i + 1][0] for i in range(2))


def test_parallel_join_waits_for_all_parents():
    log = []
    g = ComputationGraph([_node("fast", Nap(log, 0.01), x=graph_input("x")),
                          _node("slow", Nap(log, 0.08), x=graph_input("x")),
                          _node("join", Nap(log, 0.0), a="fast", b="slow")])
    run(ParallelScheduler(g).run({"x": 0}))
    assert _times(log, "start", "join") >= _times(log, "end", "slow")


def test_parallel_error_cancels_siblings():
    log = []
    g = ComputationGraph([_node("bad", Boom(), x=graph_input("x")), _node("slow", Nap(log, 2.0), x=graph_input("x")),
                          _node("sink", Plus(), a="bad", b="slow")])
    t0 = time.perf_counter()
    with pytest.raises(ZeroDivisionError):
        run(ParallelScheduler(g).run({"x": 0}))
    assert time.perf_counter() - t0 < 1.0
    assert not [1 for k, n, _ in log if k == "end" and n == "slow"]


def test_parallel_missing_inputs_and_literal_wiring_rejected():
    g = ComputationGraph([_node("a", x=graph_input("x"))])
    with pytest.raises(ValueError, match="Missing graph inputs"):
        run(ParallelScheduler(g).run({}))
    with pytest.raises(ValueError, match="unknown node"):
        ComputationGraph([_node("a", x=graph_input("x"), k=41)])     # literals are not graph wiring


def test_parallel_semaphore_zero_disables_and_default_tracks_pool():
    class P:
        size = 3
        in_process = True

        def worker_affinities(self):
            return ("worker::a-0",)

    g = ComputationGraph([_node("a", x=graph_input("x"))])
    assert ParallelScheduler(g, pool=P()).max_pending_subtasks == 24
    assert ParallelScheduler(g, pool=P(), max_pending_subtasks=5).max_pending_subtasks == 5
    assert ParallelScheduler(g).max_pending_subtasks is None
    seen = {}

    class Peek(Operator):
        name = "peek"

        def compute(self, inputs, *, context):
            seen.update(context.metadata)
            return 1

    pg = ComputationGraph([_node("p", Peek(), x=graph_input("x"))])
    run(ParallelScheduler(pg, pool=P(), max_pending_subtasks=0).run({"x": 0}))
    assert "subtask_semaphore" not in seen and seen["pool_size"] == 3 and seen["pool_in_process"] is True
    run(ParallelScheduler(pg, pool=P()).run({"x": 0}))
    assert isinstance(seen["subtask_semaphore"], asyncio.Semaphore)


def test_parallel_and_sequential_agree_on_operator_graph():
    vs = _vs(n=9, d=64)
    b = GraphBuilder()
    x = b.input("vectors")
    left = x.apply(Clipping(threshold=1.0), name="clip").apply(CoordinateWiseMedian(), name="med")
    right = x.apply(CoordinateWiseTrimmedMean(f=2), name="tm")
    g = b.build(outputs=[left.key, right.key])
    seq = run(NodeScheduler(g).run({"vectors": vs}))
    par = run(ParallelScheduler(g).run({"vectors": vs}))
    assert all(torch.equal(seq[k], par[k]) for k in ("med", "tm"))


def test_tracer_spans_cover_every_node():
    from byzpy_b200.utils.tracing import Tracer

    tr = Tracer(cuda=False)
    g = ComputationGraph([_node("a", Plus(), x=graph_input("x")), _node("b", Plus(), x="a")])
    run(ParallelScheduler(g, metadata={"tracer": tr}).run({"x": 0}))
    run(NodeScheduler(g, metadata={"tracer": tr}).run({"x": 0}))
    names = [e["name"] for e in tr.finalize()]
    assert names.count("node:a") == 2 and names.count("node:b") == 2


# -------------------------------------------------------------------------- message-aware scheduler
def test_message_delivered_before_wait_is_queued():
    async def go():
        s = MessageAwareNodeScheduler(ComputationGraph([_node("a", x=graph_input("x"))]))
        s.deliver_message("t", 1)
        s.deliver_message("t", 2)
        assert [await s.wait_for_message("t"), await s.wait_for_message("t")] == [1, 2]

    run(go())


def test_message_wakes_every_waiter_and_is_also_queued():
    async def go():
        s = MessageAwareNodeScheduler(ComputationGraph([_node("a", x=graph_input("x"))]))
        w = [asyncio.ensure_future(s.wait_for_message("t")) for _ in range(3)]
        await asyncio.sleep(0)
        s.deliver_message("t", "hi")
        assert await asyncio.gather(*w) == ["hi"] * 3
        assert await s.wait_for_message("t", timeout=0.1) == "hi"     # reference keeps the payload queued too

    run(go())


def test_message_wait_timeout_removes_waiter():
    async def go():
        s = MessageAwareNodeScheduler(ComputationGraph([_node("a", x=graph_input("x"))]))
        with pytest.raises(asyncio.TimeoutError):
            await s.wait_for_message("never", timeout=0.01)
        assert s._message_waiters.get("never") == []
        s.deliver_message("never", 5)
        assert await s.wait_for_message("never") == 5

    run(go())


def test_message_types_are_independent():
    async def go():
        s = MessageAwareNodeScheduler(ComputationGraph([_node("a", x=graph_input("x"))]))
        s.deliver_message("a", 1)
        with pytest.raises(asyncio.TimeoutError):
            await s.wait_for_message("b", timeout=0.01)
        assert await s.wait_for_message("a") == 1

    run(go())


def test_message_source_as_graph_input_and_node_input():
    g = ComputationGraph([_node("s", Plus(), a=graph_input("a"), b=MessageSource("m", field="v"))])

    async def go():
        s = MessageAwareNodeScheduler(g)
        s.deliver_message("seed", {"v": 10, "w": 0})
        s.deliver_message("m", {"v": 5})
        return await s.run({"a": MessageSource("seed", field="v")})

    assert run(go()) == {"s": 15}


def test_message_source_field_errors():
    g = ComputationGraph([_node("s", Plus(), a=MessageSource("m", field="v"))])

    async def bad(payload):
        s = MessageAwareNodeScheduler(g)
        s.deliver_message("m", payload)
        return await s.run({})

    with pytest.raises(TypeError, match="non-dict"):
        run(bad(3))
    with pytest.raises(KeyError, match="not found"):
        run(bad({"w": 1}))


def test_message_trigger_blocks_until_delivery():
    g = ComputationGraph([_node("wait", MessageTriggerOp("go")), _node("after", Plus(1), x="wait")])

    async def go():
        s = MessageAwareNodeScheduler(g)
        task = asyncio.ensure_future(s.run({}))
        await asyncio.sleep(0.02)
        assert not task.done()
        s.deliver_message("go", 41)
        return await task

    assert run(go()) == {"after": 42}


def test_message_trigger_validation():
    with pytest.raises(ValueError):
        MessageTriggerOp("")
    assert MessageTriggerOp("x").name == "message_trigger_x"
    with pytest.raises(RuntimeError, match="requires scheduler"):
        run(MessageTriggerOp("x").run({}, context=OpContext("n"), pool=None))
    g = ComputationGraph([_node("wait", MessageTriggerOp("go", timeout=0.01))])
    with pytest.raises(asyncio.TimeoutError):
        run(MessageAwareNodeScheduler(g).run({}))


# ------------------------------------------------------------------------- subtasks and dispatcher
def test_subtask_validation_and_helpers():
    with pytest.raises(TypeError):
        SubTask(fn=3)
    with pytest.raises(ValueError):
        SubTask(fn=len, max_retries=-1)
    st = SubTask(fn=divmod, args=(7, 2), name="dm")
    assert st.run() == (3, 1) and st.label == "dm" and SubTask(fn=len).label == "len"
    pinned = st.pinned_to("gpu")
    assert pinned.affinity == "gpu" and st.affinity is None and pinned.args == st.args
    assert SubTask(fn=dict, kwargs={"a": 1}).run() == {"a": 1}


@pytest.mark.parametrize("limit,size,expect", [(None, 2, 16), (0, 3, 24), (5, 2, 5), (-2, 3, 6), (-1, 0, 1), (None, 0, 1)])
def test_window_size_rules(limit, size, expect):
    assert _window_size(limit, size) == expect


class _CountingPool:
    def __init__(self, size=2, dt=0.005):
        self.size, self.dt = size, dt
        self.inflight = self.peak = self.total = 0
        self.affinities = []

    async def run_subtask(self, st):
        self.inflight += 1
        self.total += 1
        self.peak = max(self.peak, self.inflight)
        self.affinities.append(st.affinity)
        try:
            await asyncio.sleep(self.dt * (1 + (self.total % 3)))
            return st.run()
        finally:
            self.inflight -= 1

    def worker_affinities(self):
        return tuple(f"worker::w-{i}" for i in range(self.size))


def test_windowed_dispatch_bounds_inflight_and_keeps_order():
    pool = _CountingPool()
    out = run(run_subtasks_windowed(pool, (SubTask(fn=abs, args=(-i,)) for i in range(40)), 3))
    assert out == list(range(40)) and pool.peak == 3 and pool.total == 40


def test_windowed_dispatch_is_lazy_over_generators():
    pulled = []

    def gen():
        for i in range(10):
            pulled.append(i)
            yield SubTask(fn=abs, args=(i,))

    class Stall(_CountingPool):
        async def run_subtask(self, st):
            if st.args[0] == 0:
                assert len(pulled) <= 2     # only the window has been materialised so far
            return await super().run_subtask(st)

    assert run(run_subtasks_windowed(Stall(), gen(), 2)) == list(range(10))


def test_windowed_dispatch_shared_semaphore_is_released():
    async def go():
        sem = asyncio.Semaphore(2)
        pool = _CountingPool()
        out = await run_subtasks_windowed(pool, [SubTask(fn=abs, args=(i,)) for i in range(9)], 8, sem)
        assert out == list(range(9)) and pool.peak <= 2 and sem._value == 2

    run(go())


def test_windowed_dispatch_error_cancels_and_releases():
    def fail():
        raise KeyError("bad subtask")

    async def go():
        sem = asyncio.Semaphore(4)
        pool = _CountingPool(dt=0.02)
        tasks = [SubTask(fn=abs, args=(1,)), SubTask(fn=fail)] + [SubTask(fn=abs, args=(2,))] * 10
        with pytest.raises(KeyError):
            await run_subtasks_windowed(pool, tasks, 4, sem)
        await asyncio.sleep(0.1)
        assert sem._value == 4 and pool.total < len(tasks)

    run(go())


def test_windowed_dispatch_empty_iterable():
    assert run(run_subtasks_windowed(_CountingPool(), [], None)) == []


class Chunky(Operator):
    name = "chunky"
    supports_subtasks = True
    max_subtasks_inflight = 2

    def __init__(self, pieces=6):
        self.pieces = pieces
        self.computed = False

    def compute(self, inputs, *, context):
        self.computed = True
        return sum(inputs["xs"])

    def create_subtasks(self, inputs, *, context):
        xs = inputs["xs"]
        step = max(1, len(xs) // self.pieces) if self.pieces else len(xs)
        return [SubTask(fn=sum, args=(xs[i:i + step],)) for i in range(0, len(xs), step)] if self.pieces else []

    def reduce_subtasks(self, partials, inputs, *, context):
        return sum(partials)


def test_operator_uses_subtasks_only_with_multi_worker_pool():
    xs = list(range(30))
    op = Chunky()
    pool = _CountingPool(size=2)
    assert run(op.run({"xs": xs}, context=OpContext("n"), pool=pool)) == sum(xs)
    assert not op.computed and pool.total == 6 and pool.peak <= 2
    solo = _CountingPool(size=1)
    assert run(op.run({"xs": xs}, context=OpContext("n"), pool=solo)) == sum(xs) and op.computed and solo.total == 0
    op2 = Chunky()
    assert run(op2.run({"xs": xs}, context=OpContext("n"), pool=None)) == sum(xs) and op2.computed


def test_operator_falls_back_to_compute_when_no_subtasks_produced():
    op = Chunky(pieces=0)
    pool = _CountingPool(size=4)
    assert run(op.run({"xs": [1, 2, 3]}, context=OpContext("n"), pool=pool)) == 6
    assert op.computed and pool.total == 0


def test_operator_affinity_hints_round_robin_but_keep_explicit_pins():
    class Pinned(Chunky):
        def create_subtasks(self, inputs, *, context):
            sts = list(super().create_subtasks(inputs, context=context))
            sts[1] = sts[1].pinned_to("gpu")
            return sts

    pool = _CountingPool(size=2, dt=0.0)
    ctx = OpContext("n", metadata={"worker_affinities": pool.worker_affinities()})
    run(Pinned(pieces=4).run({"xs": list(range(8))}, context=ctx, pool=pool))
    assert sorted(map(str, pool.affinities)) == sorted(["worker::w-0", "gpu", "worker::w-0", "worker::w-1"])


def test_barriered_operator_takes_precedence_and_sync_result_is_accepted():
    class Barrier(Chunky):
        supports_barriered_subtasks = True

        def run_barriered_subtasks(self, inputs, *, context, pool):     # deliberately sync
            return "barriered"

    pool = _CountingPool(size=2)
    assert run(Barrier().run({"xs": [1]}, context=OpContext("n"), pool=pool)) == "barriered"
    assert run(Barrier().run({"xs": [1]}, context=OpContext("n"), pool=None)) == 1


def test_base_operator_defaults():
    op = Operator()
    with pytest.raises(NotImplementedError):
        op.compute({}, context=OpContext("n"))
    assert list(op.create_subtasks({}, context=OpContext("n"))) == []
    with pytest.raises(RuntimeError, match="reduce_subtasks"):
        op.reduce_subtasks([], {}, context=OpContext("n"))
    with pytest.raises(RuntimeError, match="barriered"):
        run(op.run_barriered_subtasks({}, context=OpContext("n"), pool=None))


# ------------------------------------------------------------------------------------ callable ops
def test_callable_op_binds_by_mapping():
    op = CallableOp(lambda a, b: a - b, input_mapping={"a": "left", "b": "right"})
    g = make_single_operator_graph(node_name="sub", operator=op, input_keys=("left", "right"))
    assert run(NodeScheduler(g).run({"left": 5, "right": 3})) == {"sub": 2}
    with pytest.raises(KeyError, match="right"):
        op.compute({"left": 1}, context=OpContext("n"))


def test_remote_callable_op_ships_one_subtask():
    op = RemoteCallableOp(pow, input_mapping={"base": "b", "exp": "e"})
    sts = list(op.create_subtasks({"b": 2, "e": 5}, context=OpContext("n")))
    assert len(sts) == 1 and sts[0].run() == 32
    pool = _CountingPool(size=2)
    assert run(op.run({"b": 2, "e": 5}, context=OpContext("n"), pool=pool)) == 32 and pool.total == 1
    with pytest.raises(RuntimeError):
        op.reduce_subtasks([], {}, context=OpContext("n"))


def test_callable_op_accepts_async_functions():
    async def double(v):
        await asyncio.sleep(0)
        return 2 * v

    g = make_single_operator_graph(node_name="d", operator=CallableOp(double, input_mapping={"v": "x"}),
                                   input_keys=("x",))
    assert run(NodeScheduler(g).run({"x": 4})) == {"d": 8}
    assert run(ParallelScheduler(g).run({"x": 4})) == {"d": 8}
