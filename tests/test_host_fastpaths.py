"""Host-side fast paths: the intra-op thread governor of the thread actor backend and the flat
(ParamArena) paths of the example nodes / P2P mixin.  CPU only."""
from __future__ import annotations

import asyncio
import os
import sys
import threading
import time

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from byzpy_b200.engine.actor.backends._local import IntraOpGovernor  # noqa: E402
from byzpy_b200.engine.actor.backends.thread import ThreadActorBackend  # noqa: E402
from byzpy_b200.engine.actor.base import ActorRef  # noqa: E402
from byzpy_b200.parallel.arena import ParamArena, flatten_grads, write_vector_to_grads_  # noqa: E402


# ------------------------------------------------------------------------------ governor
def test_governor_alone_keeps_the_full_thread_count():
    gov = IntraOpGovernor()
    gov._base = 8
    seen = {}

    def body():
        with gov:
            seen["alone"] = gov._tls.threads
    t = threading.Thread(target=body)
    t.start()
    t.join()
    assert seen["alone"] == 8
    assert gov._active == 0


def test_governor_splits_threads_between_concurrent_calls_and_restores(monkeypatch):
    calls = []
    monkeypatch.setattr(torch, "set_num_threads", lambda k: calls.append((threading.get_ident(), k)))
    gov = IntraOpGovernor()
    gov._base = 8
    inside = threading.Barrier(4)
    shares = []

    def body():
        with gov:
            inside.wait(timeout=10)          # all four are in flight at the same time
            shares.append(gov._share())
            inside.wait(timeout=10)
    ts = [threading.Thread(target=body) for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert shares == [2, 2, 2, 2]            # 8 // 4 once everybody is in
    assert gov._active == 0
    assert calls[-1][1] == 8                 # the last one out restores the process default
    assert min(k for _, k in calls) >= 1


def test_governor_remembers_the_recent_peak_then_decays():
    gov = IntraOpGovernor()
    gov._base = 8
    gov.PEAK_WINDOW = 0.05
    gov._active = 4
    assert gov._share() == 2
    gov._active = 1                          # next round's first arrival: sized by the recent peak
    assert gov._share() == 2
    time.sleep(0.08)
    assert gov._share() == 8                 # concurrency stopped: a lone call gets everything again


def test_governor_never_goes_below_one_thread_and_can_be_disabled(monkeypatch):
    gov = IntraOpGovernor()
    gov._base = 2
    gov._active = 9
    assert gov._share() == 1
    monkeypatch.setenv("BYZPY_INTRAOP_GOVERNOR", "0")
    off = IntraOpGovernor()
    assert not off.enabled
    with off:
        assert off._active == 0


def test_thread_actor_calls_run_under_the_governor():
    from byzpy_b200.engine.actor.backends import _local

    class Probe:
        def active(self):
            return _local.intra_op_governor._active

    async def main():
        ref = ActorRef(ThreadActorBackend())
        async with ref:
            await ref._backend.construct(Probe, args=(), kwargs={})
            return await ref.active()

    assert asyncio.run(main()) >= 1
    assert _local.intra_op_governor._active == 0


# ------------------------------------------------------------------------------ flat example nodes
def test_ps_example_node_matches_torch_sgd_bit_for_bit():
    from examples.ps.nodes import DistributedPSHonestNode, SmallCNN

    node = DistributedPSHonestNode(indices=list(range(256)), batch_size=16)
    torch.manual_seed(0)
    twin = SmallCNN()
    opt = torch.optim.SGD(twin.parameters(), lr=0.05, momentum=0.9)
    crit = torch.nn.CrossEntropyLoss()
    for _ in range(3):
        x, y = node.next_batch()
        g = node.local_honest_gradient(x=x, y=y)
        twin.zero_grad(set_to_none=True)
        crit(twin(x), y).backward()
        assert torch.equal(g, flatten_grads(twin))
        node.apply_server_gradient(0.5 * g)
        write_vector_to_grads_(twin, 0.5 * g)
        opt.step()
        for a, b in zip(node.model.parameters(), twin.parameters()):
            assert torch.equal(a, b)
    assert node.arena.check_bound()
    assert set(node.dump_state_dict()) == set(twin.state_dict())


def test_returned_gradient_is_a_snapshot_not_a_view_of_the_arena():
    from examples.ps.nodes import DistributedPSHonestNode

    node = DistributedPSHonestNode(indices=list(range(64)), batch_size=8)
    x, y = node.next_batch()
    g1 = node.local_honest_gradient(x=x, y=y)
    keep = g1.clone()
    x, y = node.next_batch()
    node.local_honest_gradient(x=x, y=y)
    assert torch.equal(g1, keep)


def test_p2p_mixin_flat_path_equals_the_per_parameter_path():
    from examples.p2p.nodes import P2PHonestNode

    fast = P2PHonestNode(indices=list(range(256)), seed=5)
    slow = P2PHonestNode(indices=list(range(256)), seed=5)
    slow.arena = None
    assert fast._bound_arena() is not None and slow._bound_arena() is None
    for _ in range(3):
        va, vb = fast.p2p_half_step(0.05), slow.p2p_half_step(0.05)
        assert torch.equal(va, vb)
        nb = [va + 0.01, va - 0.02, va * 1.01]
        fast.p2p_aggregate_and_set(va, nb)
        slow.p2p_aggregate_and_set(vb, nb)
        assert torch.equal(fast.get_param_vector(), slow.get_param_vector())
    assert fast._bound_arena() is not None


def test_p2p_mixin_falls_back_when_the_arena_is_unbound():
    from examples.p2p.nodes import P2PHonestNode

    node = P2PHonestNode(indices=list(range(64)), seed=1)
    node.model.zero_grad(set_to_none=True)         # user code dropped the gradient views
    assert node._bound_arena() is None
    v = node.p2p_half_step(0.05)
    assert v.numel() == node.arena.d and torch.isfinite(v).all()
    node.set_param_vector(torch.zeros_like(v))
    assert float(node.get_param_vector().abs().max()) == 0.0


def test_param_arena_on_cpu_keeps_reference_flat_layout():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    want = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    arena = ParamArena(m)
    assert torch.equal(arena.param_vector(), want)
    m(torch.randn(4, 5)).sum().backward()
    assert torch.equal(arena.grad_vector(), flatten_grads(m))
    assert arena.check_bound()


# ------------------------------------------------------------------------------ ParallelScheduler offload
def _branch_graph(op_factory, k=3):
    from byzpy_b200.engine.graph.graph import ComputationGraph, GraphNode, graph_input

    nodes = [GraphNode(f"b{i}", op_factory(i), {"x": graph_input("x")}) for i in range(k)]
    return ComputationGraph(nodes, outputs=[f"b{i}" for i in range(k)])


def test_parallel_scheduler_runs_concurrent_host_computes_on_worker_threads():
    from byzpy_b200.engine.graph.ops import CallableOp
    from byzpy_b200.engine.graph.parallel_scheduler import ParallelScheduler

    main_thread = threading.get_ident()
    seen = {}

    def make(i):
        def work(x):
            seen[i] = (threading.get_ident(), time.perf_counter())
            time.sleep(0.15)                        # a blocking, GIL-releasing compute
            return x + i
        return CallableOp(work, input_mapping={"x": "x"})

    t0 = time.perf_counter()
    out = asyncio.run(ParallelScheduler(_branch_graph(make)).run({"x": 1}))
    elapsed = time.perf_counter() - t0
    assert out == {"b0": 1, "b1": 2, "b2": 3}
    assert all(tid != main_thread for tid, _ in seen.values())      # none of them held the event loop
    assert elapsed < 0.40                                            # 3 x 0.15 s overlapped


def test_parallel_scheduler_offload_can_be_disabled_and_a_lone_node_stays_inline():
    from byzpy_b200.engine.graph.graph import ComputationGraph, GraphNode, graph_input
    from byzpy_b200.engine.graph.ops import CallableOp
    from byzpy_b200.engine.graph.parallel_scheduler import ParallelScheduler

    main_thread = threading.get_ident()
    tids = []

    def make(i):
        return CallableOp(lambda x: tids.append(threading.get_ident()) or x, input_mapping={"x": "x"})

    asyncio.run(ParallelScheduler(_branch_graph(make), metadata={"offload_host_compute": False}).run({"x": 0}))
    assert tids and all(t == main_thread for t in tids)
    tids.clear()
    lone = ComputationGraph([GraphNode("only", make(0), {"x": graph_input("x")})])
    asyncio.run(ParallelScheduler(lone).run({"x": 0}))
    assert tids == [main_thread]


def test_parallel_scheduler_offload_propagates_errors_and_async_results():
    from byzpy_b200.engine.graph.ops import CallableOp
    from byzpy_b200.engine.graph.parallel_scheduler import ParallelScheduler

    def make(i):
        if i == 1:
            def boom(x):
                raise ValueError("branch 1 failed")
            return CallableOp(boom, input_mapping={"x": "x"})

        async def later(x):
            return x * 10

        return CallableOp(lambda x: later(x), input_mapping={"x": "x"})      # compute() returns an awaitable

    with pytest.raises(ValueError, match="branch 1 failed"):
        asyncio.run(ParallelScheduler(_branch_graph(make)).run({"x": 2}))
    ok = asyncio.run(ParallelScheduler(_branch_graph(lambda i: make(0), k=2)).run({"x": 2}))
    assert ok == {"b0": 20, "b1": 20}


def test_node_scheduler_never_offloads():
    from byzpy_b200.engine.graph.ops import CallableOp
    from byzpy_b200.engine.graph.scheduler import NodeScheduler

    main_thread = threading.get_ident()
    tids = []
    make = lambda i: CallableOp(lambda x: tids.append(threading.get_ident()) or x, input_mapping={"x": "x"})  # noqa: E731
    asyncio.run(NodeScheduler(_branch_graph(make)).run({"x": 0}))
    assert tids == [main_thread] * 3


# ------------------------------------------------------------------------------ shared-memory packing
def test_register_rows_is_one_segment_equal_to_the_stack():
    from byzpy_b200.engine.storage.shared_store import cleanup_tensor, open_tensor, register_rows

    torch.manual_seed(11)
    for dtype in (torch.float32, torch.float64):
        rows = [torch.randn(1000, dtype=dtype) for _ in range(5)]
        h = register_rows(rows)
        try:
            assert h.shape == (5, 1000) and h.dtype == ("float32" if dtype == torch.float32 else "float64")
            with open_tensor(h) as arr:
                assert torch.equal(torch.from_numpy(arr.copy()), torch.stack(rows))
        finally:
            cleanup_tensor(h)
    shaped = [torch.arange(6.0).reshape(2, 3) + i for i in range(3)]          # rows are flattened
    h = register_rows(shaped)
    try:
        with open_tensor(h) as arr:
            assert arr.shape == (3, 6) and arr[2, 5] == 7.0
    finally:
        cleanup_tensor(h)


def test_attach_cached_reuses_and_bounds_its_mappings():
    from byzpy_b200.engine.storage import shared_store as ss

    handles = [ss.register_rows([torch.full((16,), float(i))]) for i in range(ss._ATTACH_LIMIT + 2)]
    try:
        first = ss.attach_cached(handles[0])
        assert first[0, 0] == 0.0
        seg = ss._ATTACHED[handles[0].name]
        del first
        assert ss.attach_cached(handles[0]) is not None and ss._ATTACHED[handles[0].name] is seg   # same mapping
        for h in handles[1:]:
            assert ss.attach_cached(h)[0, 3] == float(handles.index(h))
        assert len(ss._ATTACHED) <= ss._ATTACH_LIMIT and handles[0].name not in ss._ATTACHED       # oldest evicted
        ss.cleanup_tensor(handles[-1])                                   # unlinked name, cached mapping still readable
        assert ss.attach_cached(handles[-1])[0, 0] == float(len(handles) - 1)
    finally:
        for h in handles:
            ss.cleanup_tensor(h)


def _threads_in_worker():
    import torch as _t

    return _t.get_num_threads()


@pytest.mark.real_actor_backends
def test_process_pool_workers_split_the_cores(monkeypatch):
    from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig
    from byzpy_b200.engine.graph.subtask import SubTask

    monkeypatch.setattr(os, "cpu_count", lambda: 8)

    async def main():
        pool = ActorPool([ActorPoolConfig(backend="process", count=2)])
        await pool.start()
        try:
            return [await pool.run_subtask(SubTask(fn=_threads_in_worker, affinity=a))
                    for a in pool.worker_affinities()]
        finally:
            await pool.shutdown()

    assert asyncio.run(main()) == [4, 4]


class _ThreadProbe:
    def threads(self):
        return torch.get_num_threads()


@pytest.mark.real_actor_backends
def test_process_actors_share_follows_the_number_of_live_actors(monkeypatch):
    from byzpy_b200.engine.actor.backends import process as proc

    monkeypatch.setattr(os, "cpu_count", lambda: 8)
    for b in list(proc._LIVE):                    # leftovers of other tests do not count
        b._closed = True

    async def main():
        a, b = proc.ProcessActorBackend(), proc.ProcessActorBackend()
        for be in (a, b):
            await be.start()
            await be.construct(_ThreadProbe, args=(), kwargs={})
        two = [await a.call("threads"), await b.call("threads")]
        await b.close()
        one = await a.call("threads")             # b is gone: a gets the whole machine on its next call
        await a.close()
        return two, one

    two, one = asyncio.run(main())
    assert two == [4, 4] and one == 8
    monkeypatch.setenv("BYZPY_INTRAOP_GOVERNOR", "0")
    assert proc._thread_share() == 0


def test_wrap_payload_batches_equal_tensors_into_one_segment_and_round_trips():
    from byzpy_b200.engine.actor import ipc

    torch.manual_seed(12)
    grads = [torch.randn(3, 7000) for _ in range(5)]
    wrapped = ipc.wrap_payload({"honest_grads": grads, "k": 3, "pair": (grads[0], torch.ones(2))})
    assert wrapped["honest_grads"][0] == ipc._SHM_BATCH_MARK            # one segment for the five gradients
    assert wrapped["pair"][0][0] == ipc._SHM_MARK and wrapped["k"] == 3
    assert ipc.wrap_payload(wrapped["honest_grads"]) is wrapped["honest_grads"]   # idempotent
    back = ipc.unwrap_payload(wrapped)
    assert isinstance(back["honest_grads"], list) and len(back["honest_grads"]) == 5
    for a, b in zip(back["honest_grads"], grads):
        assert a.shape == (3, 7000) and torch.equal(a, b)
    assert torch.equal(back["pair"][0], grads[0]) and isinstance(back["pair"], tuple)
    tup = ipc.unwrap_payload(ipc.wrap_payload(tuple(grads)))
    assert isinstance(tup, tuple) and torch.equal(tup[4], grads[4])
    # not batched: mixed dtypes / shapes, tiny tensors, single element
    for payload in ([grads[0], grads[1].double()], [torch.ones(2), torch.ones(3)], [torch.ones(4), torch.ones(4)], [grads[0]]):
        w = ipc.wrap_payload(payload)
        assert isinstance(w, list) and all(x[0] == ipc._SHM_MARK for x in w)
        for a, b in zip(ipc.unwrap_payload(w), payload):
            assert torch.equal(a, b) and a.dtype == b.dtype
