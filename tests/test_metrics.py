"""The metrics registry (``utils/metrics.py``) and what the framework reports into it."""
import asyncio
import logging
import urllib.request

import pytest
import torch

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
from byzpy_b200.engine.graph.ops import make_single_operator_graph
from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig
from byzpy_b200.engine.graph.scheduler import NodeScheduler
from byzpy_b200.engine.graph.subtask import SubTask
from byzpy_b200.engine.parameter_server.ps import ParameterServer
from byzpy_b200.utils import metrics


@pytest.fixture
def collecting():
    metrics.reset()
    metrics.enable()
    yield metrics
    metrics.enable(False)
    metrics.reset()


def test_registry_is_off_by_default_and_costs_nothing():
    metrics.reset()
    assert not metrics.enabled()
    metrics.inc("x")
    metrics.observe("y", 1.0)
    metrics.set_gauge("z", 3)
    with metrics.timer("t"):
        pass
    assert metrics.snapshot() == {} and metrics.to_prometheus_text() == "\n"


def test_counters_gauges_histograms_and_the_text_exposition(collecting):
    m = collecting
    m.REGISTRY.describe("jobs_total", "jobs finished")
    m.inc("jobs_total")
    m.inc("jobs_total", 2)
    m.inc("errors_total", labels={"kind": "io"})
    m.inc("errors_total", 3, labels={"kind": "net"})
    m.set_gauge("queue_depth", 7)
    for v in (0.0004, 0.02, 0.02, 3.0, 99.0):
        m.observe("latency_seconds", v)
    with m.timer("block_seconds"):
        pass
    snap = m.snapshot()
    assert snap["jobs_total"] == 3 and snap["errors_total"] == {"kind=io": 1, "kind=net": 3} and snap["queue_depth"] == 7
    assert snap["latency_seconds"]["count"] == 5 and abs(snap["latency_seconds"]["sum"] - 102.0404) < 1e-9
    assert snap["block_seconds"]["count"] == 1
    text = m.to_prometheus_text()
    assert "# HELP jobs_total jobs finished\n# TYPE jobs_total counter\njobs_total 3\n" in text
    assert 'errors_total{kind="net"} 3' in text and "# TYPE queue_depth gauge\nqueue_depth 7" in text
    # cumulative buckets: 1 observation <= 0.5 ms, 3 <= 25 ms, 4 <= 5 s, 5 in total
    assert 'latency_seconds_bucket{le="0.0005"} 1' in text and 'latency_seconds_bucket{le="0.025"} 3' in text
    assert 'latency_seconds_bucket{le="5"} 4' in text and 'latency_seconds_bucket{le="+Inf"} 5' in text
    assert "latency_seconds_count 5" in text
    m.reset()
    assert m.snapshot() == {}


def test_http_endpoint_serves_the_exposition(collecting):
    collecting.inc("served_total", 4)
    srv = collecting.serve(port=0)
    try:
        port = srv.server_address[1]
        body = urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=5).read().decode()
        assert "served_total 4" in body
        with pytest.raises(Exception):
            urllib.request.urlopen(f"http://127.0.0.1:{port}/nope", timeout=5)
    finally:
        srv.shutdown()
        srv.server_close()


class _Honest:
    def __init__(self, g, fail_from=None):
        self.g, self.calls, self.fail_from = torch.tensor(g), 0, fail_from

    def honest_gradient_for_next_batch(self):
        self.calls += 1
        if self.fail_from is not None and self.calls >= self.fail_from:
            raise RuntimeError("disk on fire")
        return self.g

    def apply_server_gradient(self, g):
        pass


def test_parameter_server_pool_and_dispatch_report_into_the_registry(collecting, caplog, monkeypatch, tmp_path):
    monkeypatch.setenv("BYZPY_POOL_DISPATCH", "reference")
    nodes = [_Honest([1.0, 2.0]), _Honest([2.0, 1.0]), _Honest([3.0, 3.0], fail_from=2)]
    ps = ParameterServer(nodes, [], CoordinateWiseMedian(), tolerate_failures=True, fused=False)
    with caplog.at_level(logging.WARNING, logger="byzpy_b200"):
        for _ in range(3):
            asyncio.run(ps.round())
    snap = collecting.snapshot()
    assert snap["byzpy_ps_rounds_total"] == {"path=generic": 3}
    assert snap["byzpy_ps_round_seconds"]["count"] == 3
    assert snap["byzpy_ps_node_failures_total"] == {"kind=honest": 2}
    assert sum("honest node 2 skipped" in r.getMessage() for r in caplog.records) == 2

    async def through_pool():
        pool = ActorPool([ActorPoolConfig("thread", count=2)])
        await pool.start()
        try:
            graph = make_single_operator_graph(node_name="agg", operator=CoordinateWiseMedian(chunk_size=64),
                                               input_keys=("gradients",))
            await NodeScheduler(graph, pool=pool).run({"gradients": [torch.randn(1000) for _ in range(5)]})
            marker = str(tmp_path / "attempted")     # (the function travels by value: state must live outside it)

            def sometimes():
                import os

                if not os.path.exists(marker):
                    open(marker, "w").close()
                    raise RuntimeError("first attempt fails")
                return 7

            assert await pool.run_subtask(SubTask(fn=sometimes, max_retries=1)) == 7
        finally:
            await pool.shutdown()

    asyncio.run(through_pool())
    snap = collecting.snapshot()
    assert snap["byzpy_pool_subtasks_total"] >= 3 and snap["byzpy_pool_subtask_retries_total"] == 1
    assert snap["byzpy_operator_runs_total"] == {"op=coordinate-wise-median,route=pool": 1}
