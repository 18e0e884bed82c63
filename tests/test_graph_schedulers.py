"""Scheduling layer: graph validation, lazy builder, schedulers, sessions, executor."""
import asyncio
import time

import pytest
import torch

from byzpy_b200 import OperatorExecutor, run_operator
from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
from byzpy_b200.attacks import EmpireAttack
from byzpy_b200.engine.graph.graph import ComputationGraph, GraphInput, GraphNode, graph_input
from byzpy_b200.engine.graph.lazy import GraphBuilder
from byzpy_b200.engine.graph.operator import MessageTriggerOp, OpContext, Operator
from byzpy_b200.engine.graph.ops import CallableOp, RemoteCallableOp, make_single_operator_graph
from byzpy_b200.engine.graph.parallel_scheduler import ParallelScheduler
from byzpy_b200.engine.graph.scheduler import MessageAwareNodeScheduler, MessageSource, NodeScheduler
from byzpy_b200.engine.graph.session import ExecutionSession
from byzpy_b200.engine.graph.subtask import SubTask
from byzpy_b200.pre_aggregators import Clipping


def run(coro):
    return asyncio.run(coro)


class Add(Operator):
    name = "add"

    def __init__(self, k=1):
        self.k = k

    def compute(self, inputs, *, context):
        return sum(inputs.values()) + self.k


class Sleepy(Operator):
    name = "sleepy"

    def __init__(self, log, dt=0.05):
        self.log, self.dt = log, dt

    async def compute(self, inputs, *, context):
        self.log.append(("start", context.node_name, time.perf_counter()))
        await asyncio.sleep(self.dt)
        self.log.append(("end", context.node_name, time.perf_counter()))
        return context.node_name


def test_graph_validation_and_order():
    a = GraphNode("a", Add(), {"x": graph_input("x")})
    b = GraphNode("b", Add(), {"x": "a"})
    c = GraphNode("c", Add(), {"x": "a", "y": "b"})
    g = ComputationGraph([c, a, b])
    order = [n.name for n in g.nodes_in_order()]
    assert order.index("a") < order.index("b") < order.index("c")
    assert g.outputs == ["c"] and g.required_inputs == frozenset({"x"})
    with pytest.raises(ValueError):
        ComputationGraph([])
    with pytest.raises(ValueError):
        ComputationGraph([a, GraphNode("a", Add(), {})])
    with pytest.raises(ValueError):
        ComputationGraph([GraphNode("z", Add(), {"x": "missing"})])
    with pytest.raises(ValueError):
        ComputationGraph([GraphNode("p", Add(), {"x": "q"}), GraphNode("q", Add(), {"x": "p"})])
    with pytest.raises(ValueError):
        ComputationGraph([a], outputs=["nope"])
    src = GraphInput.from_message("grad", field="v")
    g2 = ComputationGraph([GraphNode("m", Add(), {"x": src})])
    assert g2.required_inputs == frozenset()


def test_node_scheduler_runs_and_validates_inputs():
    g = ComputationGraph([GraphNode("a", Add(1), {"x": graph_input("x")}), GraphNode("b", Add(10), {"x": "a"})])
    assert run(NodeScheduler(g).run({"x": 1})) == {"b": 12}
    with pytest.raises(ValueError):
        run(NodeScheduler(g).run({}))


def test_metadata_reaches_operator():
    seen = {}

    class Peek(Operator):
        def compute(self, inputs, *, context):
            seen.update(context.metadata)
            seen["node"] = context.node_name
            return 0

    g = make_single_operator_graph(node_name="p", operator=Peek(), input_keys=("x",))
    run(NodeScheduler(g, metadata={"tag": 7}).run({"x": 0}))
    assert seen["tag"] == 7 and seen["node"] == "p"


def test_parallel_scheduler_runs_branches_concurrently_and_respects_limit():
    log = []
    nodes = [GraphNode(f"s{i}", Sleepy(log), {"x": graph_input("x")}) for i in range(4)]
    nodes.append(GraphNode("join", Add(0), {f"i{i}": f"s{i}" for i in range(4)}) if False else
                 GraphNode("join", CallableOp(lambda **kw: sorted(kw.values()), input_mapping={f"i{i}": f"i{i}" for i in range(4)}),
                           {f"i{i}": f"s{i}" for i in range(4)}))
    g = ComputationGraph(nodes, outputs=["join"])
    t0 = time.perf_counter()
    out = run(ParallelScheduler(g).run({"x": 0}))
    dt = time.perf_counter() - t0
    assert out["join"] == ["s0", "s1", "s2", "s3"] and dt < 0.15
    log.clear()
    t0 = time.perf_counter()
    run(ParallelScheduler(g, max_concurrent_nodes=1).run({"x": 0}))
    assert time.perf_counter() - t0 >= 0.19


def test_parallel_scheduler_is_dataflow_not_waves():
    # chain a->b is short, c is long and independent: b must start before c ends
    log = []
    g = ComputationGraph([
        GraphNode("a", Sleepy(log, 0.02), {"x": graph_input("x")}),
        GraphNode("b", Sleepy(log, 0.02), {"x": "a"}),
        GraphNode("c", Sleepy(log, 0.15), {"x": graph_input("x")}),
    ], outputs=["b", "c"])
    run(ParallelScheduler(g).run({"x": 0}))
    t = {(k, n): ts for k, n, ts in log}
    assert t[("start", "b")] < t[("end", "c")]


def test_parallel_scheduler_error_propagates_and_semaphore_defaults():
    class Boom(Operator):
        def compute(self, inputs, *, context):
            raise RuntimeError("boom")

    g = ComputationGraph([GraphNode("x1", Boom(), {"x": graph_input("x")}), GraphNode("x2", Add(), {"x": graph_input("x")})],
                         outputs=["x1"])
    with pytest.raises(RuntimeError):
        run(ParallelScheduler(g).run({"x": 1}))

    class FakePool:
        size = 3

        def worker_affinities(self):
            return ()

    assert ParallelScheduler(g, pool=FakePool()).max_pending_subtasks == 24
    assert ParallelScheduler(g, pool=FakePool(), max_pending_subtasks=0).max_pending_subtasks == 0
    assert ParallelScheduler(g).max_pending_subtasks is None


def test_operator_dispatch_order_and_window():
    calls = []

    class Pool:
        size = 2

        async def run_subtask(self, st):
            calls.append(st.name)
            await asyncio.sleep(0.001)
            return st.fn(*st.args)

    class Chunky(Operator):
        supports_subtasks = True
        max_subtasks_inflight = 3

        def compute(self, inputs, *, context):
            return "direct"

        def create_subtasks(self, inputs, *, context):
            return [SubTask(fn=lambda i=i: i * i, name=f"t{i}") for i in range(7)]

        def reduce_subtasks(self, partials, inputs, *, context):
            return list(partials)

    c = OpContext("n", {})
    assert run(Chunky().run({}, context=c, pool=None)) == "direct"
    assert run(Chunky().run({}, context=c, pool=Pool())) == [i * i for i in range(7)]  # submission order kept

    class One:
        size = 1

    assert run(Chunky().run({}, context=c, pool=One())) == "direct"  # needs pool.size > 1

    class Barriered(Chunky):
        supports_barriered_subtasks = True

        async def run_barriered_subtasks(self, inputs, *, context, pool):
            return "barriered"

    assert run(Barriered().run({}, context=c, pool=One())) == "barriered"  # no size test on this path


def test_worker_affinity_round_robin():
    seen = []

    class Pool:
        size = 2

        async def run_subtask(self, st):
            seen.append(st.affinity)
            return 0

    class Op(Operator):
        supports_subtasks = True

        def create_subtasks(self, inputs, *, context):
            return [SubTask(fn=int), SubTask(fn=int, affinity="gpu"), SubTask(fn=int)]

        def reduce_subtasks(self, partials, inputs, *, context):
            return len(partials)

    ctx = OpContext("n", {"worker_affinities": ("worker::a-0", "worker::a-1")})
    assert run(Op().run({}, context=ctx, pool=Pool())) == 3
    assert seen == ["worker::a-0", "gpu", "worker::a-0"]


def test_message_aware_scheduler():
    async def scenario():
        g = ComputationGraph([GraphNode("n", CallableOp(lambda v: v * 2, input_mapping={"v": "v"}),
                                        {"v": MessageSource("grad", field="vec")})])
        s = MessageAwareNodeScheduler(g)
        task = asyncio.ensure_future(s.run({}))
        await asyncio.sleep(0.01)
        s.deliver_message("grad", {"vec": 21})
        assert (await task) == {"n": 42}
        # the delivered payload is also cached for a later waiter (reference behaviour)
        assert await s.wait_for_message("grad") == {"vec": 21}
        with pytest.raises(asyncio.TimeoutError):
            await s.wait_for_message("other", timeout=0.01)
        s.deliver_message("x", "not-a-dict")
        g2 = ComputationGraph([GraphNode("n", Add(), {"x": MessageSource("x", field="f")})])
        s2 = MessageAwareNodeScheduler(g2)
        s2.deliver_message("x", "not-a-dict")
        with pytest.raises(TypeError):
            await s2.run({})
        # graph input given as a MessageSource
        g3 = make_single_operator_graph(node_name="n", operator=Add(0), input_keys=("x",))
        s3 = MessageAwareNodeScheduler(g3)
        s3.deliver_message("m", 5)
        assert await s3.run({"x": MessageSource("m")}) == {"n": 5}
        # trigger op
        g4 = ComputationGraph([GraphNode("t", MessageTriggerOp("go", timeout=1.0), {})])
        s4 = MessageAwareNodeScheduler(g4)
        s4.deliver_message("go", "now")
        assert await s4.run({}) == {"t": "now"}
        with pytest.raises(RuntimeError):
            await NodeScheduler(g4).run({})

    run(scenario())
    with pytest.raises(ValueError):
        MessageTriggerOp("")


def test_lazy_builder():
    b = GraphBuilder()
    x = b.input("vectors")
    assert b.input("vectors").key == "vectors"
    clipped = x.apply(Clipping(threshold=1.0))
    agg = clipped.apply(CoordinateWiseMedian(), name="final")
    assert clipped.key == "pre-agg/clipping_0" and agg.key == "final"
    g = b.build(outputs=[agg.key])
    vs = [torch.randn(10) * 5 for _ in range(5)]
    out = run(NodeScheduler(g).run({"vectors": vs}))["final"]
    exp = CoordinateWiseMedian().aggregate(Clipping(threshold=1.0).pre_aggregate(vs))
    assert torch.allclose(out, exp)
    with pytest.raises(TypeError):
        x.apply("nope")
    with pytest.raises(ValueError):
        GraphBuilder().build(outputs=["a"])
    with pytest.raises(ValueError):
        b.build(outputs=["missing"])
    extra = x.apply(Add(), input_key="a", extra_inputs={"b": clipped, "c": "final"}, name="mix")
    assert b._nodes["mix"].inputs["b"] == clipped.key and isinstance(b._nodes["mix"].inputs["a"], GraphInput)


def test_execution_session_cache_and_future():
    count = {"n": 0}

    class Count(Operator):
        name = "count"

        def compute(self, inputs, *, context):
            count["n"] += 1
            return inputs["x"] + 1

    g = ComputationGraph([GraphNode("a", Count(), {"x": graph_input("x")}), GraphNode("b", Count(), {"x": "a"})])

    async def scenario():
        async with ExecutionSession() as s:
            assert await s.execute(g, {"x": 1}) == {"b": 3}
            assert count["n"] == 2 and s.is_cached("a") and s.get_cached("b") == 3
            assert await s.execute(g, {"x": 100}) == {"b": 3}  # cache is keyed by node name
            assert count["n"] == 2
            g2 = ComputationGraph([GraphNode("a", Count(), {"x": graph_input("x")}), GraphNode("c", Count(), {"x": "a"})])
            assert await s.execute(g2, {"x": 1}) == {"c": 3} and count["n"] == 3
            s.clear_cache()
            assert not s.is_cached("a")
            with pytest.raises(KeyError):
                s.get_cached("a")
            fut = s.execute_async(g, {"x": 5})
            assert fut.output_keys == ("b",)
            assert (await fut) == {"b": 7} and fut.done()
            with pytest.raises(RuntimeError):
                fut.result()
        assert not s._result_cache
        s2 = ExecutionSession(cache_intermediate=False)
        assert await s2.execute(g, {"x": 0}) == {"b": 2} and not s2._result_cache

    run(scenario())


def test_executor_and_run_operator():
    vs = [torch.randn(20) for _ in range(5)]
    exp = CoordinateWiseMedian().aggregate(vs)
    assert torch.equal(run(run_operator(CoordinateWiseMedian(), {"gradients": vs})), exp)
    assert torch.equal(run(OperatorExecutor(CoordinateWiseMedian(), input_keys=("vecs",)).run({"vecs": vs})), exp)
    with pytest.raises(ValueError):
        OperatorExecutor(EmpireAttack())
    with pytest.raises(ValueError):
        OperatorExecutor(Add())
    with pytest.raises(TypeError):
        OperatorExecutor(object())
    out = run(run_operator(EmpireAttack(), {"honest_grads": vs}, input_keys=("honest_grads",)))
    assert torch.allclose(out, -torch.stack(vs).mean(0), atol=1e-6)


def test_remote_callable_op_without_pool_runs_locally():
    op = RemoteCallableOp(lambda a, b: a + b, input_mapping={"a": "p", "b": "q"})
    g = make_single_operator_graph(node_name="r", operator=op, input_keys=("p", "q"))
    assert run(NodeScheduler(g).run({"p": 1, "q": 2})) == {"r": 3}
    with pytest.raises(KeyError):
        CallableOp(lambda a: a, input_mapping={"a": "zz"}).compute({}, context=OpContext("n"))
