"""Adaptive pool dispatch (engine/graph/operator.py): the reference's rule is where a workload starts, the
measured faster route is where it ends up."""
import asyncio
import time

import pytest
import torch

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
from byzpy_b200.aggregators.geometric_wise import GeometricMedian
from byzpy_b200.engine.graph.operator import OpContext, Operator
from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig
from byzpy_b200.engine.graph.subtask import SubTask


def _sleepy(seconds, value):
    time.sleep(seconds)
    return value


class Probe(Operator):
    """compute() takes ``direct_s``; each of its two subtasks takes ``sub_s``."""

    name = "probe"
    supports_subtasks = True
    input_key = "xs"

    def __init__(self, direct_s, sub_s):
        self.direct_s, self.sub_s = direct_s, sub_s
        self.routes = []

    def compute(self, inputs, *, context):
        self.routes.append("direct")
        time.sleep(self.direct_s)
        return sum(float(x.sum()) for x in inputs["xs"])

    def create_subtasks(self, inputs, *, context):
        self.routes.append("pool")
        xs = inputs["xs"]
        h = len(xs) // 2
        return [SubTask(fn=_sleepy, args=(self.sub_s, sum(float(x.sum()) for x in part)), kwargs={})
                for part in (xs[:h], xs[h:])]

    def reduce_subtasks(self, partials, inputs, *, context):
        return sum(partials)


async def _drive(op, calls, xs):
    pool = ActorPool([ActorPoolConfig(backend="thread", count=2)])
    await pool.start()
    try:
        ctx = OpContext(node_name="n")
        return [await op.run({"xs": xs}, context=ctx, pool=pool) for _ in range(calls)]
    finally:
        await pool.shutdown()


def test_exploration_order_then_the_faster_route(monkeypatch):
    monkeypatch.delenv("BYZPY_POOL_DISPATCH", raising=False)
    xs = [torch.ones(8) for _ in range(4)]
    slow_pool = Probe(direct_s=0.0, sub_s=0.05)
    outs = asyncio.run(_drive(slow_pool, 9, xs))
    assert slow_pool.routes[:4] == ["pool", "direct", "pool", "direct"]          # both routes, cold and warm
    assert set(slow_pool.routes[4:]) == {"direct"} and len(set(outs)) == 1 and outs[0] == 32.0
    rep = next(iter(slow_pool.dispatch_report().values()))
    assert rep["direct"] < rep["pool"] and rep["n_pool"] == 2 and rep["n_direct"] == 7

    slow_direct = Probe(direct_s=0.08, sub_s=0.0)
    asyncio.run(_drive(slow_direct, 8, xs))
    assert slow_direct.routes[:4] == ["pool", "direct", "pool", "direct"] and set(slow_direct.routes[4:]) == {"pool"}


def test_the_slower_route_is_retried_periodically(monkeypatch):
    monkeypatch.delenv("BYZPY_POOL_DISPATCH", raising=False)
    monkeypatch.setattr(Operator, "_REEXPLORE", 6)
    op = Probe(direct_s=0.0, sub_s=0.02)
    asyncio.run(_drive(op, 13, [torch.ones(4) for _ in range(4)]))
    assert op.routes[4:] == ["direct", "pool", "direct", "direct", "direct", "direct", "direct", "pool", "direct"]


def test_a_new_shape_starts_from_the_reference_rule_again(monkeypatch):
    monkeypatch.delenv("BYZPY_POOL_DISPATCH", raising=False)
    op = Probe(direct_s=0.0, sub_s=0.03)

    async def go():
        pool = ActorPool([ActorPoolConfig(backend="thread", count=2)])
        await pool.start()
        try:
            ctx = OpContext(node_name="n")
            for _ in range(5):
                await op.run({"xs": [torch.ones(8)] * 4}, context=ctx, pool=pool)
            await op.run({"xs": [torch.ones(16)] * 4}, context=ctx, pool=pool)
        finally:
            await pool.shutdown()

    asyncio.run(go())
    assert op.routes[4] == "direct" and op.routes[5] == "pool" and len(op.dispatch_report()) == 2


def test_fixed_rule_on_request(monkeypatch):
    monkeypatch.setenv("BYZPY_POOL_DISPATCH", "reference")
    op = Probe(direct_s=0.0, sub_s=0.02)
    asyncio.run(_drive(op, 5, [torch.ones(4) for _ in range(4)]))
    assert op.routes == ["pool"] * 5 and op.dispatch_report() == {}


@pytest.mark.parametrize("mk", [CoordinateWiseMedian, GeometricMedian])
def test_both_routes_agree_for_real_aggregators(monkeypatch, mk):
    """Subtask route, barriered route (GeometricMedian) and direct route compute the same aggregate."""
    monkeypatch.delenv("BYZPY_POOL_DISPATCH", raising=False)
    torch.manual_seed(0)
    grads = [torch.randn(5000) for _ in range(9)]
    op = mk()

    async def go():
        pool = ActorPool([ActorPoolConfig(backend="thread", count=2)])
        await pool.start()
        try:
            ctx = OpContext(node_name="agg")
            return [await op.run({"gradients": grads}, context=ctx, pool=pool) for _ in range(4)]
        finally:
            await pool.shutdown()

    outs = asyncio.run(go())
    rep = next(iter(op.dispatch_report().values()))
    assert rep["n_pool"] == 2 and rep["n_direct"] == 2
    for o in outs[1:]:
        torch.testing.assert_close(o, outs[0], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(outs[0], mk().aggregate(grads), rtol=1e-4, atol=1e-5)
