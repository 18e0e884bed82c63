"""Fine-grained tiers of the worker pool and actor plumbing, one behaviour per test: worker
acquisition / parking / release order, retries, shutdown, function shipping caches, pool channels,
the channel router, endpoints, shared-memory payloads and the thread backend's serial-execution
guarantee (mirrors reference tests/engine/graph/test_pool.py and tests/engine/actor/*)."""
import asyncio
import pickle
import threading
import time

import cloudpickle
import numpy as np
import pytest
import torch

from byzpy_b200.engine.actor.base import ActorBackend, ActorRef
from byzpy_b200.engine.actor.channels import ChannelRef, Endpoint, open_channel
from byzpy_b200.engine.actor.ipc import unwrap_payload, wrap_payload
from byzpy_b200.engine.actor.router import BackendRecord, ChannelRouter, channel_router
from byzpy_b200.engine.actor.backends.thread import ThreadActorBackend
from byzpy_b200.engine.graph import pool as pool_mod
from byzpy_b200.engine.graph.pool import (ActorPool, ActorPoolChannel, ActorPoolConfig, _PoolWorker,
                                           _SubTaskWorker)
from byzpy_b200.engine.graph.subtask import SubTask
from byzpy_b200.engine.storage.shared_store import (SharedTensorHandle, cleanup_tensor, is_handle,
                                                     materialize, open_tensor, register_tensor)


def run(coro):
    return asyncio.run(coro)


class Fake:
    """In-line backend with a per-call delay; records which worker object ran what."""

    made = []

    def __init__(self, delay=0.01, fail_first=0):
        self.delay, self.fail_first = delay, fail_first
        self.obj, self.calls, self.closed, self.started = None, 0, False, 0
        self.mail = {}
        Fake.made.append(self)

    async def start(self):
        self.started += 1

    async def construct(self, cls, *, args, kwargs):
        self.obj = cls(*args, **kwargs)

    async def call(self, method, *args, **kwargs):
        self.calls += 1
        if self.calls <= self.fail_first:
            raise RuntimeError("injected failure")
        await asyncio.sleep(self.delay)
        return getattr(self.obj, method)(*args, **kwargs)

    async def close(self):
        self.closed = True

    async def get_endpoint(self):
        return Endpoint("fake", "", f"f{id(self)}")

    async def chan_open(self, name):
        self.mail.setdefault(name, asyncio.Queue())
        return await self.get_endpoint()

    async def chan_put(self, *, from_ep, to_ep, name, payload):
        target = [b for b in Fake.made if f"f{id(b)}" == to_ep.actor_id][0]
        target.mail.setdefault(name, asyncio.Queue()).put_nowait(payload)

    async def chan_get(self, *, ep, name, timeout):
        try:
            return await asyncio.wait_for(self.mail[name].get(), timeout)
        except asyncio.TimeoutError:
            return None


@pytest.fixture
def fake(monkeypatch):
    Fake.made.clear()
    monkeypatch.setattr(pool_mod, "resolve_backend", lambda spec: spec if not isinstance(spec, str) else Fake())
    return Fake


def ident(x):
    return x


def whoami():
    return threading.get_ident()


# ---------------------------------------------------------------------------------- pool lifecycle
def test_size_before_and_after_start(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=2), ActorPoolConfig("thread", count=3)])
        assert pool.size == 5 and pool.worker_affinities() == ()
        await pool.start()
        assert pool.size == 5 and len(pool.worker_affinities()) == 5
        await pool.shutdown()
        assert pool.size == 5 and pool.worker_affinities() == ()

    run(go())


def test_start_is_idempotent_and_lazy_on_first_subtask(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=2)])
        assert await pool.run_subtask(SubTask(ident, (3,))) == 3      # started on demand
        await pool.start()
        await pool.start()
        assert len(fake.made) == 2 and all(b.started == 1 for b in fake.made)
        await pool.shutdown()

    run(go())


def test_default_and_custom_worker_names(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=1), ActorPoolConfig("thread", count=2, name="agg")])
        await pool.start()
        assert [w.name for w in pool._workers] == ["actor-0", "agg-0", "agg-1"]
        assert pool.worker_affinities() == ("worker::actor-0", "worker::agg-0", "worker::agg-1")
        assert all("cpu" in w.capabilities and f"worker::{w.name}" in w.capabilities for w in pool._workers)
        await pool.shutdown()

    run(go())


def test_empty_pool_refuses_work(fake):
    with pytest.raises(RuntimeError, match="no workers"):
        run(ActorPool([]).run_subtask(SubTask(ident, (1,))))


def test_run_many_empty_and_order(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=3)])
        assert await pool.run_many([]) == []
        assert await pool.run_many([SubTask(ident, (i,)) for i in range(11)]) == list(range(11))
        await pool.shutdown()

    run(go())


def test_concurrency_equals_worker_count(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=4)])
        await pool.start()
        for b in fake.made:
            b.delay = 0.05
        t0 = time.perf_counter()
        await pool.run_many([SubTask(ident, (i,)) for i in range(8)])
        dt = time.perf_counter() - t0
        assert 0.09 < dt < 0.35                           # two waves of four (one at a time would take 0.4 s)
        assert sorted(b.calls for b in fake.made) == [2, 2, 2, 2]
        await pool.shutdown()

    run(go())


def test_waiters_are_parked_and_served_fifo(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=1)])
        await pool.start()
        order = []

        async def job(i):
            await pool.run_subtask(SubTask(ident, (i,)))
            order.append(i)

        tasks = [asyncio.ensure_future(job(i)) for i in range(5)]
        await asyncio.sleep(0.005)
        assert len(pool._waiting[None]) == 4 and pool._available.empty()
        await asyncio.gather(*tasks)
        assert order == list(range(5)) and pool._available.qsize() == 1
        await pool.shutdown()

    run(go())


def test_release_prefers_capability_specific_waiter(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("gpu", count=1, name="g")])
        await pool.start()
        order = []

        async def job(tag, affinity):
            await pool.run_subtask(SubTask(ident, (tag,), affinity=affinity))
            order.append(tag)

        first = asyncio.ensure_future(job("running", None))
        await asyncio.sleep(0.002)
        anyone = asyncio.ensure_future(job("anyone", None))
        await asyncio.sleep(0.002)
        gpu = asyncio.ensure_future(job("gpu", "gpu"))
        await asyncio.gather(first, anyone, gpu)
        assert order == ["running", "gpu", "anyone"]
        await pool.shutdown()

    run(go())


def test_affinity_skips_busy_incapable_workers(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=2, name="c"), ActorPoolConfig("gpu", count=1, name="g")])
        await pool.start()
        out = await asyncio.gather(*[pool.run_subtask(SubTask(ident, (i,), affinity="cpu")) for i in range(6)])
        assert out == list(range(6))
        gpu_backend = [w for w in pool._workers if "gpu" in w.capabilities][0].backend
        assert gpu_backend.calls == 0 and pool._available.qsize() == 3
        await pool.shutdown()

    run(go())


def test_unknown_affinity_raises_and_leaves_pool_usable(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=2)])
        await pool.start()
        with pytest.raises(RuntimeError, match="No actor in the pool"):
            await pool.run_subtask(SubTask(ident, (1,), affinity="worker::nobody-9"))
        assert pool._available.qsize() == 2
        assert await pool.run_subtask(SubTask(ident, (2,))) == 2
        await pool.shutdown()

    run(go())


def test_retry_budget_is_per_subtask_and_worker_is_released(fake):
    async def go():
        flaky = Fake(delay=0.0, fail_first=3)
        pool = ActorPool([ActorPoolConfig(flaky, count=1)])
        await pool.start()
        with pytest.raises(RuntimeError, match="injected"):
            await pool.run_subtask(SubTask(ident, (1,), max_retries=1))      # attempts 1-2 fail
        assert pool._available.qsize() == 1
        assert await pool.run_subtask(SubTask(ident, (7,), max_retries=1)) == 7   # attempt 3 fails, 4 succeeds
        assert flaky.calls == 4
        await pool.shutdown()

    run(go())


def test_exception_from_subtask_propagates_with_type(fake):
    def bad():
        raise LookupError("inside worker")

    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=1)])
        with pytest.raises(LookupError, match="inside worker"):
            await pool.run_subtask(SubTask(bad))
        assert await pool.run_subtask(SubTask(ident, (1,))) == 1
        await pool.shutdown()

    run(go())


def test_shutdown_fails_parked_waiters_and_closes_backends(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=1)])
        await pool.start()
        fake.made[0].delay = 0.2
        running = asyncio.ensure_future(pool.run_subtask(SubTask(ident, (1,))))
        await asyncio.sleep(0.005)
        parked = asyncio.ensure_future(pool.run_subtask(SubTask(ident, (2,))))
        await asyncio.sleep(0.005)
        await pool.shutdown()
        with pytest.raises(RuntimeError, match="shutdown"):
            await parked
        running.cancel()
        assert fake.made[0].closed and not pool._waiting

    run(go())


def test_pool_restarts_after_shutdown(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=2)])
        await pool.start()
        await pool.shutdown()
        assert await pool.run_many([SubTask(ident, (i,)) for i in range(3)]) == [0, 1, 2]
        assert len(fake.made) == 4
        await pool.shutdown()

    run(go())


def test_pool_pickles_as_configuration_only(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=2, name="w")])
        await pool.start()
        clone = pickle.loads(pickle.dumps(pool))
        assert clone.size == 2 and not clone._started and clone._workers == []
        assert clone.configs == pool.configs
        await pool.shutdown()

    run(go())


@pytest.mark.parametrize("backends,expect", [(["thread"], True), (["thread", "gpu"], True), (["gpu:0"], True),
                                             (["process"], False), (["thread", "tcp://h:1"], False)])
def test_in_process_property(backends, expect):
    assert ActorPool([ActorPoolConfig(b) for b in backends]).in_process is expect


def test_in_process_is_false_for_backend_instances():
    assert ActorPool([ActorPoolConfig(ThreadActorBackend())]).in_process is False


# ------------------------------------------------------------------------------- function shipping
def test_pool_worker_caches_serialised_functions(fake, monkeypatch):
    dumps = []
    real = cloudpickle.dumps
    monkeypatch.setattr(pool_mod.cloudpickle, "dumps", lambda fn: dumps.append(fn) or real(fn))
    w = _PoolWorker(backend=Fake(), capabilities={"cpu"}, name="w")
    b1, b2 = w._serialized_fn(ident), w._serialized_fn(ident)
    assert b1 is b2 and dumps == [ident]
    w._serialized_fn(whoami)
    assert list(w._fn_cache) == [ident, whoami]
    w._serialized_fn(ident)
    assert list(w._fn_cache) == [whoami, ident]           # LRU order refreshed


def test_pool_worker_function_cache_is_bounded(fake):
    w = _PoolWorker(backend=Fake(), capabilities={"cpu"}, name="w")
    w._fn_cache_limit = 3
    fns = [(lambda k: (lambda: k))(i) for i in range(5)]
    for f in fns:
        w._serialized_fn(f)
    assert list(w._fn_cache) == fns[2:]


def test_pool_worker_accepts_unhashable_callables(fake):
    class Unhashable:
        __hash__ = None

        def __call__(self, x):
            return x + 1

    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=1)])
        assert await pool.run_subtask(SubTask(Unhashable(), (1,))) == 2
        await pool.shutdown()

    run(go())


def test_subtask_worker_unpickles_once_per_payload():
    w = _SubTaskWorker()
    blob = cloudpickle.dumps(ident)
    assert w.execute(blob, (4,), {}) == 4 and w.execute(blob, (5,), {}) == 5
    assert len(w._fns) == 1
    first = next(iter(w._fns.values()))
    w.execute(blob, (6,), {})
    assert next(iter(w._fns.values())) is first


def test_subtask_worker_cache_evicts_oldest():
    w = _SubTaskWorker()
    w._CACHE_LIMIT = 2
    blobs = [cloudpickle.dumps((lambda k: (lambda: k))(i)) for i in range(3)]
    assert [w.execute(b, (), {}) for b in blobs] == [0, 1, 2]
    assert list(w._fns) == blobs[1:]
    assert w.execute(blobs[0], (), {}) == 0 and list(w._fns) == [blobs[2], blobs[0]]


def test_closures_and_kwargs_are_shipped_by_value(fake):
    k = 10

    def scaled(x, *, by=1):
        return (x + k) * by

    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=2)])
        assert await pool.run_subtask(SubTask(scaled, (1,), {"by": 3})) == 33
        await pool.shutdown()

    run(go())


# ------------------------------------------------------------------------------------ pool channels
def test_pool_channel_send_recv_and_cache(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=3, name="w")])
        ch = await pool.open_channel("partials")
        assert isinstance(ch, ActorPoolChannel) and ch.name == "partials" and ch.workers == ("w-0", "w-1", "w-2")
        assert await pool.open_channel("partials") is ch
        assert await pool.open_channel("other") is not ch
        await ch.send("w-0", "w-2", {"v": 1})
        await ch.send("w-1", "w-2", {"v": 2})
        assert [(await ch.recv("w-2", timeout=1))["v"] for _ in range(2)] == [1, 2]
        assert await ch.recv("w-0", timeout=0.01) is None
        assert ch.endpoint("w-1") == await pool._workers[1].endpoint()
        await pool.shutdown()
        assert pool._channel_cache == {}

    run(go())


def test_pool_channel_unknown_worker_errors(fake):
    async def go():
        pool = ActorPool([ActorPoolConfig("thread", count=1, name="w")])
        ch = await pool.open_channel("c")
        with pytest.raises(KeyError, match="No channel bound"):
            ch.channel("w-9")
        with pytest.raises(KeyError, match="No endpoint known"):
            ch.endpoint("w-9")
        with pytest.raises(KeyError):
            await ch.send("w-0", "w-9", 1)
        await pool.shutdown()

    run(go())


# ---------------------------------------------------------------------- refs / endpoints / router
def test_endpoint_fields_str_and_remote_flag():
    ep = Endpoint("tcp", "10.0.0.1:5000", "a1")
    assert (ep.scheme, ep.address, ep.actor_id) == ("tcp", "10.0.0.1:5000", "a1")
    assert str(ep) == "tcp://10.0.0.1:5000/a1" and str(Endpoint("thread", "", "x")) == "thread:/x"
    assert ep.is_remote() and Endpoint("ucx", "h:1", "a").is_remote() and not Endpoint("process", "", "a").is_remote()
    assert ep == Endpoint("tcp", "10.0.0.1:5000", "a1") and hash(ep) == hash(tuple(ep))
    assert pickle.loads(pickle.dumps(ep)) == ep


def test_actor_ref_forwards_calls_but_not_dunders():
    calls = []

    class Rec:
        async def call(self, method, *a, **k):
            calls.append((method, a, k))
            return len(calls)

    ref = ActorRef(Rec())
    assert run(ref.step(1, lr=0.1)) == 1 and calls == [("step", (1,), {"lr": 0.1})]
    assert ref.anything.__name__ == "anything"
    with pytest.raises(AttributeError):
        ref.__deepcopy__
    assert not hasattr(ref, "__wrapped_missing__")


def test_actor_ref_context_manager_starts_and_closes(fake):
    be = Fake()

    async def go():
        async with ActorRef(be) as ref:
            assert be.started == 1 and not be.closed
            ch = await ref.open_channel("m")
            assert isinstance(ch, ChannelRef) and ch.name == "m" and ch.endpoint == await ref.endpoint()
            assert "m" in repr(ch)
            same = await open_channel(be, "m")
            assert same.endpoint == ch.endpoint
        assert be.closed

    run(go())


def test_backends_satisfy_the_structural_protocol():
    assert isinstance(ThreadActorBackend(), ActorBackend)
    assert isinstance(Fake(), ActorBackend)
    assert not isinstance(object(), ActorBackend)


def test_channel_router_register_resolve_unregister():
    r = ChannelRouter()
    a, b = object(), object()
    r.register("thread", "x", a)
    r.register("process", "x", b)
    assert r.resolve("thread", "x") is a and r.resolve("process", "x") is b and r.resolve("gpu", "x") is None
    r.register("thread", "x", b)                         # re-registration replaces
    assert r.resolve("thread", "x") is b
    r.unregister("thread", "x")
    r.unregister("thread", "x")                          # idempotent
    assert r.resolve("thread", "x") is None
    rec = BackendRecord("s", "i", a)
    assert (rec.scheme, rec.actor_id, rec.backend) == ("s", "i", a)
    assert isinstance(channel_router, ChannelRouter)


def test_thread_backend_registers_with_global_router_until_closed():
    async def go():
        be = ThreadActorBackend()
        await be.start()
        await be.construct(dict, args=(), kwargs={})
        ep = await be.chan_open("c")
        assert channel_router.resolve(ep.scheme, ep.actor_id) is be
        await be.close()
        assert channel_router.resolve(ep.scheme, ep.actor_id) is None

    run(go())


def test_thread_actor_methods_never_overlap_and_stay_on_one_thread():
    class Probe:
        def __init__(self):
            self.active = self.peak = 0
            self.threads = set()

        def work(self):
            self.active += 1
            self.peak = max(self.peak, self.active)
            self.threads.add(threading.get_ident())
            time.sleep(0.01)
            self.active -= 1
            return self.peak

        def stats(self):
            return self.peak, len(self.threads)

    async def go():
        be = ThreadActorBackend()
        async with ActorRef(be) as ref:
            await be.construct(Probe, args=(), kwargs={})
            await asyncio.gather(*[ref.work() for _ in range(8)])
            assert await ref.stats() == (1, 1)

    run(go())


def test_thread_actor_call_before_construct_fails():
    async def go():
        be = ThreadActorBackend()
        await be.start()
        with pytest.raises(RuntimeError, match="not constructed"):
            await be.call("x")
        await be.close()

    run(go())


def test_thread_channel_timeout_returns_none_and_fifo_order():
    async def go():
        a, b = ThreadActorBackend(), ThreadActorBackend()
        for be in (a, b):
            await be.start()
            await be.construct(dict, args=(), kwargs={})
        ca, cb = await ActorRef(a).open_channel("q"), await ActorRef(b).open_channel("q")
        assert await cb.recv(timeout=0.01) is None
        for i in range(5):
            await ca.send(cb.endpoint, i)
        assert [await cb.recv(timeout=1) for _ in range(5)] == list(range(5))
        other = await ActorRef(b).open_channel("other")
        assert await other.recv(timeout=0.01) is None          # mailboxes are per name
        await a.close()
        await b.close()

    run(go())


# -------------------------------------------------------------------------- shared-memory payloads
def test_shared_store_register_open_cleanup():
    arr = np.arange(12, dtype=np.float32).reshape(3, 4)
    h = register_tensor(arr)
    assert isinstance(h, SharedTensorHandle) and h.shape == (3, 4) and h.dtype == "float32" and is_handle(h)
    with open_tensor(h) as view:
        assert np.array_equal(view, arr)
        view[0, 0] = 99                                  # a live mapping, not a copy
    with open_tensor({"name": h.name, "shape": [3, 4], "dtype": "float32"}) as view:
        assert view[0, 0] == 99
    assert torch.equal(materialize(h)[1], torch.tensor([4.0, 5.0, 6.0, 7.0]))
    cleanup_tensor(h)
    cleanup_tensor(h)                                    # second unlink is a no-op
    with pytest.raises(FileNotFoundError):
        with open_tensor(h):
            pass


def test_shared_store_handles_tensors_empty_and_noncontiguous():
    t = torch.arange(20.0).reshape(4, 5).t()             # non-contiguous
    h = register_tensor(t)
    assert torch.equal(materialize(h), t.contiguous())
    cleanup_tensor(h)
    e = register_tensor(np.zeros((0, 3), dtype=np.int64))
    assert materialize(e).shape == (0, 3)
    cleanup_tensor(e)
    assert not is_handle({"name": "x"}) and is_handle({"name": "x", "shape": (1,), "dtype": "f4", "extra": 1})
    with pytest.raises(TypeError):
        cleanup_tensor(42)


def test_materialize_accepts_plain_values():
    assert torch.equal(materialize([1, 2]), torch.tensor([1, 2]))
    assert torch.equal(materialize(np.ones(2)), torch.ones(2, dtype=torch.float64))
    t = torch.ones(3)
    assert materialize(t) is t


def test_wrap_payload_is_idempotent_and_preserves_structure():
    ep = Endpoint("thread", "", "a")
    payload = {"t": torch.arange(3.0), "n": np.arange(2), "nested": [(torch.ones(1), "s"), {"k": 1}], "ep": ep,
               "scalar": 1.5, "none": None}
    w = wrap_payload(payload)
    assert wrap_payload(w["t"]) is w["t"]                                   # already wrapped
    assert w["ep"] == ep and isinstance(w["ep"], Endpoint)
    assert isinstance(w["nested"], list) and isinstance(w["nested"][0], tuple)
    back = unwrap_payload(w)
    assert torch.equal(back["t"], payload["t"]) and torch.equal(back["n"], torch.arange(2))
    assert back["nested"][0][1] == "s" and back["nested"][1] == {"k": 1}
    assert back["scalar"] == 1.5 and back["none"] is None and back["ep"] == ep


def test_unwrap_consumes_the_segment():
    w = wrap_payload(torch.ones(4))
    handle = w[1]
    assert torch.equal(unwrap_payload(w), torch.ones(4))
    with pytest.raises(FileNotFoundError):
        with open_tensor(handle):
            pass


def test_wrap_detaches_grad_tensors_and_keeps_dtype():
    x = torch.ones(3, dtype=torch.float64, requires_grad=True)
    back = unwrap_payload(wrap_payload([x * 2]))[0]
    assert back.dtype == torch.float64 and not back.requires_grad and torch.equal(back, torch.full((3,), 2.0, dtype=torch.float64))


def test_a_waiter_cancelled_right_after_being_handed_a_worker_gives_it_back():
    """Lost wake-up: ``_release`` resolves a waiter's future, the waiter is cancelled before it resumes (the
    windowed runner does that when a sibling subtask fails) -- the worker must return to the pool."""
    async def scenario():
        pool = ActorPool([ActorPoolConfig(backend="thread", count=1)])
        await pool.start()
        try:
            held = await pool._acquire(None)                     # the only worker is busy
            waiter = asyncio.ensure_future(pool._acquire(None))
            await asyncio.sleep(0)                               # the waiter is parked on its future
            await pool._release(held)                            # hands the worker to the waiter's future ...
            waiter.cancel()                                      # ... and the waiter dies before it resumes
            with pytest.raises(asyncio.CancelledError):
                await waiter
            got = await asyncio.wait_for(pool._acquire(None), timeout=1.0)
            assert got is held
            await pool._release(got)
            assert await asyncio.wait_for(pool.run_subtask(SubTask(fn=_add, args=(2, 3), kwargs={})), timeout=5.0) == 5
        finally:
            await pool.shutdown()

    asyncio.run(scenario())


def _add(a, b):
    return a + b


def test_gpu_actor_results_are_recorded_on_the_callers_stream():
    """The helper behind GPUActorBackend._on_stream, on stand-ins (no GPU here): every CUDA tensor of a nested
    result is recorded on the consuming stream, host tensors and other objects are left alone."""
    from byzpy_b200.engine.actor.backends.gpu import _record_stream

    class FakeTensor:
        def __init__(self, cuda):
            self.is_cuda, self.recorded = cuda, []

        def record_stream(self, s):
            self.recorded.append(s)

    a, b, c = FakeTensor(True), FakeTensor(False), FakeTensor(True)
    _record_stream({"x": [a, (b, "text", 3)], "y": c, "z": None}, "caller-stream")
    assert a.recorded == ["caller-stream"] and c.recorded == ["caller-stream"] and b.recorded == []
    _record_stream(torch.ones(3), "s")       # a real CPU tensor: nothing to do, nothing raised
