"""ActorPool logic (fake backend) and real actor backends (thread / process / gpu / tcp) incl. the
cross-backend channel matrix."""
import asyncio
import socket
import threading

import pytest
import torch

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean
from byzpy_b200.aggregators.geometric_wise import GeometricMedian, MultiKrum
from byzpy_b200.configs.actor import set_actor
from byzpy_b200.engine.actor.backends.gpu import UCXRemoteActorBackend
from byzpy_b200.engine.actor.backends.remote import RemoteActorBackend, RemoteActorServer
from byzpy_b200.engine.actor.backends.thread import ThreadActorBackend
from byzpy_b200.engine.actor.base import ActorRef
from byzpy_b200.engine.actor.channels import Endpoint
from byzpy_b200.engine.actor.factory import resolve_backend
from byzpy_b200.engine.actor.ipc import unwrap_payload, wrap_payload
from byzpy_b200.engine.graph import pool as pool_mod
from byzpy_b200.engine.graph.ops import make_single_operator_graph
from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig
from byzpy_b200.engine.graph.scheduler import NodeScheduler
from byzpy_b200.engine.graph.subtask import SubTask


def run(coro):
    return asyncio.run(coro)


class _FakeBackend:
    """Runs the worker object inline; records calls (reference test technique)."""

    instances = []

    def __init__(self, fail_first=0):
        self.obj = None
        self.calls = 0
        self.fail_first = fail_first
        self.closed = False
        _FakeBackend.instances.append(self)

    async def start(self):
        pass

    async def construct(self, cls, *, args, kwargs):
        self.obj = cls(*args, **kwargs)

    async def call(self, method, *args, **kwargs):
        self.calls += 1
        if self.calls <= self.fail_first:
            raise RuntimeError("injected failure")
        await asyncio.sleep(0.005)
        return getattr(self.obj, method)(*args, **kwargs)

    async def close(self):
        self.closed = True

    async def get_endpoint(self):
        return Endpoint("fake", "", str(id(self)))

    async def chan_open(self, name):
        return await self.get_endpoint()


@pytest.fixture
def fake_backends(monkeypatch):
    _FakeBackend.instances.clear()
    monkeypatch.setattr(pool_mod, "resolve_backend", lambda spec: spec if not isinstance(spec, str) else _FakeBackend())
    return _FakeBackend


def sq(x):
    return x * x


def test_pool_size_capabilities_and_affinity(fake_backends):
    async def scenario():
        pool = ActorPool([ActorPoolConfig("thread", count=2, name="cpu"), ActorPoolConfig("gpu", count=1, name="g")])
        assert pool.size == 3
        await pool.start()
        assert pool.worker_affinities() == ("worker::cpu-0", "worker::cpu-1", "worker::g-0")
        assert await pool.run_many([SubTask(sq, (i,)) for i in range(6)]) == [i * i for i in range(6)]
        gpu_worker = [w for w in pool._workers if "gpu" in w.capabilities][0]
        before = gpu_worker.backend.calls
        await asyncio.gather(*[pool.run_subtask(SubTask(sq, (3,), affinity="gpu")) for _ in range(4)])
        assert gpu_worker.backend.calls == before + 4
        assert await pool.run_subtask(SubTask(sq, (5,), affinity="worker::cpu-1")) == 25
        with pytest.raises(RuntimeError, match="No actor in the pool"):
            await pool.run_subtask(SubTask(sq, (1,), affinity="tpu"))
        await pool.shutdown()
        assert all(b.closed for b in fake_backends.instances)

    run(scenario())


def test_pool_retries(fake_backends):
    async def scenario():
        be = _FakeBackend(fail_first=2)
        pool = ActorPool([ActorPoolConfig(be, count=1)])
        await pool.start()
        assert await pool.run_subtask(SubTask(sq, (4,), max_retries=2)) == 16
        be.calls, be.fail_first = 0, 5
        with pytest.raises(RuntimeError):
            await pool.run_subtask(SubTask(sq, (4,), max_retries=1))
        await pool.shutdown()

    run(scenario())


def test_explicit_capabilities_and_infer():
    assert ActorPoolConfig("gpu").resolved_capabilities() == ("gpu",)
    assert ActorPoolConfig("ucx://h:1").resolved_capabilities() == ("gpu",)
    assert ActorPoolConfig("thread").resolved_capabilities() == ("cpu",)
    assert ActorPoolConfig("thread", capabilities=("x",)).resolved_capabilities() == ("x",)
    assert isinstance(resolve_backend("tcp://127.0.0.1:1"), RemoteActorBackend)
    assert isinstance(resolve_backend("ucx://127.0.0.1:1"), UCXRemoteActorBackend)
    assert isinstance(set_actor("thread"), ThreadActorBackend)
    with pytest.raises(ValueError):
        resolve_backend("carrier-pigeon")


class Counter:
    def __init__(self, start=0):
        self.v = start

    def add(self, k):
        self.v += k
        return self.v

    async def aadd(self, k):
        self.v += k
        return self.v

    def echo(self, t):
        return t * 2


@pytest.mark.parametrize("spec", ["thread", "gpu", "process"])
def test_actor_backends_call_and_tensors(spec):
    async def scenario():
        be = resolve_backend(spec)
        async with ActorRef(be) as ref:
            await be.construct(Counter, args=(10,), kwargs={})
            assert await ref.add(5) == 15
            assert await ref.aadd(1) == 16
            out = await ref.echo(torch.arange(4.0))
            assert torch.equal(out, torch.arange(4.0) * 2)
            with pytest.raises(Exception):
                await ref.nope()

    run(scenario())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _ServerThread:
    """Actor server on a background thread with its own event loop."""

    def __init__(self, cls=RemoteActorServer):
        self.port = _free_port()
        self.server = cls("127.0.0.1", self.port)
        self.loop = asyncio.new_event_loop()
        self.ready = threading.Event()
        self.thread = threading.Thread(target=self._main, daemon=True)
        self.thread.start()
        assert self.ready.wait(5)

    def _main(self):
        asyncio.set_event_loop(self.loop)
        self.loop.run_until_complete(self.server.start())
        self.ready.set()
        self.loop.run_forever()

    def stop(self):
        fut = asyncio.run_coroutine_threadsafe(self.server.stop(), self.loop)
        fut.result(5)
        self.loop.call_soon_threadsafe(self.loop.stop)
        self.thread.join(5)


def test_tcp_actor_server_roundtrip():
    srv = _ServerThread()
    try:
        async def scenario():
            be = resolve_backend(f"tcp://127.0.0.1:{srv.port}")
            async with ActorRef(be) as ref:
                await be.construct(Counter, args=(), kwargs={"start": 2})
                assert await ref.add(3) == 5
                assert torch.equal(await ref.echo(torch.ones(3)), torch.full((3,), 2.0))
                ep = await ref.endpoint()
                assert ep.scheme == "tcp" and ep.address.endswith(str(srv.port))
                with pytest.raises(RuntimeError):
                    await ref.missing_method()

        run(scenario())
    finally:
        srv.stop()


class _Slow:
    async def slow(self, seconds, tag):
        await asyncio.sleep(seconds)
        return tag

    def fast(self, tag):
        return tag


class _SlowTensor:
    def big(self, seconds):
        import time

        time.sleep(seconds)
        return torch.ones(1 << 18)          # 1 MB: travels back through a shared-memory segment

    def ping(self):
        return "pong"


def test_process_actor_reply_of_a_timed_out_call_does_not_leak_shared_memory():
    import glob

    async def scenario():
        be = resolve_backend("process")
        async with ActorRef(be) as ref:
            await be.construct(_SlowTensor, args=(), kwargs={})
            assert await ref.ping() == "pong"
            before = set(glob.glob("/dev/shm/psm_*"))
            with pytest.raises(asyncio.TimeoutError):
                await asyncio.wait_for(ref.big(0.3), timeout=0.05)
            assert await ref.ping() == "pong"          # queued behind the slow call, answered with ITS reply
            await asyncio.sleep(0.2)
            assert set(glob.glob("/dev/shm/psm_*")) <= before, "the abandoned reply's segment was not released"
            out = await ref.big(0.0)
            assert out.shape == (1 << 18,) and float(out.sum()) == float(1 << 18)

    run(scenario())


def test_tcp_actor_call_cancelled_by_a_timeout_does_not_poison_the_next_call():
    """``ParameterServer(node_timeout=...)`` wraps node calls in ``wait_for``: a call that times out leaves its
    reply in flight, and the next call on the same connection must not receive it."""
    srv = _ServerThread()
    try:
        async def scenario():
            be = resolve_backend(f"tcp://127.0.0.1:{srv.port}")
            async with ActorRef(be) as ref:
                await be.construct(_Slow, args=(), kwargs={})
                with pytest.raises(asyncio.TimeoutError):
                    await asyncio.wait_for(ref.slow(0.4, "late"), timeout=0.05)
                assert await ref.fast("mine") == "mine"            # reconnected: not the late reply
                await asyncio.sleep(0.5)
                assert await ref.fast("again") == "again"
                assert await ref.slow(0.01, "ok") == "ok"

        run(scenario())
    finally:
        srv.stop()


def test_cross_backend_channel_matrix():
    srv = _ServerThread()
    try:
        async def scenario():
            specs = ["thread", "gpu", "process", f"tcp://127.0.0.1:{srv.port}"]
            backends = []
            for s in specs:
                be = resolve_backend(s)
                await be.start()
                await be.construct(Counter, args=(), kwargs={})
                backends.append(be)
            refs = [ActorRef(b) for b in backends]
            chans = [await r.open_channel("grads") for r in refs]
            eps = [await r.endpoint() for r in refs]
            for i, src in enumerate(chans):
                for j, dst in enumerate(eps):
                    if i == j:
                        continue
                    payload = {"from": i, "t": torch.full((2,), float(10 * i + j))}
                    await src.send(dst, payload)
                    got = await chans[j].recv(timeout=2.0)
                    assert got["from"] == i and torch.equal(got["t"], payload["t"]), (specs[i], specs[j])
            assert await chans[0].recv(timeout=0.05) is None
            for b in backends:
                await b.close()

        run(scenario())
    finally:
        srv.stop()


def test_ipc_wrap_unwrap_roundtrip():
    payload = {"a": torch.arange(6.0).reshape(2, 3), "b": [torch.ones(2), 7], "c": ("x", torch.zeros(1))}
    back = unwrap_payload(wrap_payload(payload))
    assert torch.equal(back["a"], payload["a"]) and torch.equal(back["b"][0], payload["b"][0]) and back["b"][1] == 7
    assert isinstance(back["c"], tuple) and torch.equal(back["c"][1], torch.zeros(1))


@pytest.mark.real_actor_backends
@pytest.mark.parametrize("backend", ["thread", "gpu", "process"])
def test_operators_through_real_pools(backend):
    async def scenario():
        g = [torch.randn(500) for _ in range(10)]
        pool = ActorPool([ActorPoolConfig(backend=backend, count=3)])
        await pool.start()
        try:
            for mk in (lambda: CoordinateWiseMedian(chunk_size=64), lambda: CoordinateWiseTrimmedMean(f=2, chunk_size=64),
                       lambda: MultiKrum(f=2, q=3), lambda: GeometricMedian()):
                graph = make_single_operator_graph(node_name="agg", operator=mk(), input_keys=("gradients",))
                out = (await NodeScheduler(graph, pool=pool).run({"gradients": g}))["agg"]
                assert torch.allclose(out, mk().aggregate(g), rtol=1e-5, atol=1e-6)
            ch = await pool.open_channel("c")
            a, b = ch.workers[0], ch.workers[1]
            await ch.send(a, b, {"v": torch.ones(3)})
            assert torch.equal((await ch.recv(b, timeout=2.0))["v"], torch.ones(3))
            assert (await pool.open_channel("c")) is ch
            with pytest.raises(KeyError):
                ch.channel("nobody")
        finally:
            await pool.shutdown()

    run(scenario())


def test_ucx_transport_pools_endpoints_locks_per_peer_and_retries_once():
    """The ``ucx://`` data plane (transports/ucx.py): one pooled endpoint per (host, port) reused across
    exchanges, a per-peer lock, one transparent retry after the connection dropped, and namedtuple
    payloads (Endpoint) surviving the by-value codec (reference transports/ucx.py:110-133)."""
    from byzpy_b200.engine.actor.backends.gpu import UCXRemoteActorServer
    from byzpy_b200.engine.actor.channels import Endpoint
    from byzpy_b200.engine.actor.transports import ucx as ucx_t

    srv = _ServerThread(UCXRemoteActorServer)
    try:
        async def scenario():
            be = resolve_backend(f"ucx://127.0.0.1:{srv.port}")
            await be.start()
            await be.construct(Counter, args=(), kwargs={})
            ep = await be.chan_open("box")
            assert ep.scheme == "ucx"
            await ucx_t.chan_put("127.0.0.1", srv.port, ep.actor_id, "box", {"ep": ep, "t": torch.ones(3)})
            first = await ucx_t.get_endpoint("127.0.0.1", srv.port)
            got = await ucx_t.chan_get("127.0.0.1", srv.port, ep.actor_id, "box", 2.0)
            assert isinstance(got["ep"], Endpoint) and got["ep"] == ep and torch.equal(got["t"], torch.ones(3))
            assert (await ucx_t.get_endpoint("127.0.0.1", srv.port)) is first          # pooled, not re-dialled
            # concurrent exchanges with one peer are serialised by its lock (no interleaved frames)
            await asyncio.gather(*[ucx_t.chan_put("127.0.0.1", srv.port, ep.actor_id, "box", i) for i in range(20)])
            seen = sorted([await ucx_t.chan_get("127.0.0.1", srv.port, ep.actor_id, "box", 2.0) for _ in range(20)])
            assert seen == list(range(20))
            first[1].close()                                                          # the connection drops ...
            await ucx_t.chan_put("127.0.0.1", srv.port, ep.actor_id, "box", "again")   # ... one retry re-dials
            assert await ucx_t.chan_get("127.0.0.1", srv.port, ep.actor_id, "box", 2.0) == "again"
            assert (await ucx_t.get_endpoint("127.0.0.1", srv.port)) is not first
            assert await ucx_t.chan_get("127.0.0.1", srv.port, ep.actor_id, "box", 0.05) is None
            # a CLIENT-side timeout leaves the server's reply in flight: the endpoint must not be reused, or the
            # next exchange would read that stale reply as its own
            before = await ucx_t.get_endpoint("127.0.0.1", srv.port)
            with pytest.raises(asyncio.TimeoutError):
                await ucx_t.request("127.0.0.1", srv.port, {"op": "chan_get", "actor_id": ep.actor_id, "name": "idle",
                                                            "timeout": 0.5}, timeout=0.05)
            assert (await ucx_t.get_endpoint("127.0.0.1", srv.port)) is not before
            await ucx_t.chan_put("127.0.0.1", srv.port, ep.actor_id, "box", "fresh")
            assert await ucx_t.chan_get("127.0.0.1", srv.port, ep.actor_id, "box", 2.0) == "fresh"
            await ucx_t.clear_pool()
            await be.close()

        run(scenario())
    finally:
        srv.stop()


def test_channels_between_local_actors_and_a_ucx_actor_server():
    """Mailboxes across the ``ucx://`` scheme from both sides: a thread actor posts to an actor hosted by a
    ``UCXRemoteActorServer`` and reads that actor's mailbox remotely; the hosted actor answers through its own backend.
    On a CPU-only box the CUDA-IPC codec carries host copies, the routing is the same."""
    from byzpy_b200.engine.actor.backends.gpu import UCXRemoteActorServer
    from byzpy_b200.engine.actor.transports import ucx as ucx_t

    srv = _ServerThread(UCXRemoteActorServer)
    try:
        async def scenario():
            local = resolve_backend("thread")
            remote = resolve_backend(f"ucx://127.0.0.1:{srv.port}")
            for be in (local, remote):
                await be.start()
                await be.construct(Counter, args=(), kwargs={})
            lref, rref = ActorRef(local), ActorRef(remote)
            lch, rch = await lref.open_channel("m"), await rref.open_channel("m")
            lep, rep = await lref.endpoint(), await rref.endpoint()
            assert rep.scheme == "ucx" and rep.is_remote() and not lep.is_remote()
            # local -> remote mailbox, read back by the remote actor's own handle
            await lch.send(rep, {"k": 1, "t": torch.arange(4.0)})
            got = await rch.recv(timeout=2.0)
            assert got["k"] == 1 and torch.equal(got["t"], torch.arange(4.0))
            # ... and read REMOTELY by the local backend (chan_get on a foreign endpoint)
            await lch.send(rep, "second")
            assert await local.chan_get(ep=rep, name="m", timeout=2.0) == "second"
            assert await local.chan_get(ep=rep, name="m", timeout=0.05) is None
            # an endpoint nobody can route to
            from byzpy_b200.engine.actor.channels import Endpoint

            with pytest.raises(RuntimeError):
                await local.chan_put(from_ep=lep, to_ep=Endpoint("carrier-pigeon", "", "x"), name="m", payload=1)
            with pytest.raises(RuntimeError):
                await local.chan_get(ep=Endpoint("carrier-pigeon", "", "x"), name="m", timeout=0.01)
            await ucx_t.clear_pool()
            for be in (local, remote):
                await be.close()

        run(scenario())
    finally:
        srv.stop()
