"""Remaining host-side tiers: data helpers, array backends, dependency report, legacy transports /
runner / cluster, the async parameter-server facade, tracing, the import alias, attack subtask
paths and the CLI (mirrors reference tests/engine/test_node_runner.py, test_node_cluster.py,
tests/engine/transport/*, tests/test_cli.py, tests/configs/*)."""
import asyncio
import json
import queue
import socket
import struct
import sys
import time

import numpy as np
import pytest
import torch

from byzpy_b200 import _dependencies
from byzpy_b200.attacks import EmpireAttack, GaussianAttack, InfAttack, LittleAttack, MimicAttack, SignFlipAttack
from byzpy_b200.cli import _load_subclasses, build_parser, main as cli_main
from byzpy_b200.configs.backend import get_backend, set_backend, use_backend
from byzpy_b200.engine.backend.ndarray import get_array_backend
from byzpy_b200.engine.graph.operator import OpContext
from byzpy_b200.engine.node_cluster import NodeCluster
from byzpy_b200.engine.node_runner import NodeRunner
from byzpy_b200.engine.parameter_server.decentralized import DecentralizedParameterServer
from byzpy_b200.engine.parameter_server.runner import ParameterServerRunner, mean_aggregate
from byzpy_b200.engine.transport.local import LocalTransport
from byzpy_b200.engine.transport.tcp import TcpTransport
from byzpy_b200.engine.transport.tcp_simple import TcpMailbox, send_message
from byzpy_b200.utils.data import batch_source, evaluate, mnist_like, shard_indices
from byzpy_b200.utils.tracing import Tracer, cuda_time_ms, nvtx_range


def run(coro):
    return asyncio.run(coro)


# --------------------------------------------------------------------------------------- data helpers
def test_mnist_like_shapes_ranges_and_determinism(tmp_path):
    x, y = mnist_like(64, root=str(tmp_path))
    assert x.shape == (64, 1, 28, 28) and x.dtype == torch.float32 and y.shape == (64,) and y.dtype == torch.int64
    assert 0.0 <= x.min() and x.max() <= 1.0 and 0 <= y.min() and y.max() <= 9
    x2, y2 = mnist_like(64, root=str(tmp_path))
    assert torch.equal(x, x2) and torch.equal(y, y2)
    xt, yt = mnist_like(64, train=False, root=str(tmp_path))
    assert not torch.equal(y, yt)                                   # held-out split differs
    xs, _ = mnist_like(64, seed=9, root=str(tmp_path))
    assert not torch.equal(x, xs)


def test_synthetic_mnist_is_learnable(tmp_path):
    x, y = mnist_like(1500, root=str(tmp_path))
    model = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(784, 10))
    opt = torch.optim.SGD(model.parameters(), lr=0.5)
    for _ in range(30):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
    xt, yt = mnist_like(500, train=False, root=str(tmp_path))
    loss, acc = evaluate(model, xt, yt, "cpu", batch=128)
    assert acc > 0.8 and loss < 1.5 and model.training


@pytest.mark.parametrize("n,k", [(10, 3), (7, 7), (3, 5), (0, 2)])
def test_shard_indices_partition_with_stride(n, k):
    shards = shard_indices(n, k)
    assert len(shards) == k and sorted(i for s in shards for i in s) == list(range(n))
    assert all(s == list(range(j, n, k)) for j, s in enumerate(shards))


def test_batch_source_epochs_reshuffle_without_repeats():
    x, y = torch.arange(10.0).unsqueeze(1), torch.arange(10)
    nxt = batch_source(x, y, 4, seed=1)
    b1, b2 = nxt(), nxt()
    seen = torch.cat([b1[1], b2[1]])
    assert len(set(seen.tolist())) == 8 and torch.equal(b1[0].squeeze(1).long(), b1[1])
    b3 = nxt()                                                      # 2 left < 4: new epoch
    assert b3[1].numel() == 4
    again = batch_source(x, y, 4, seed=1)
    assert torch.equal(again()[1], b1[1])


# -------------------------------------------------------------------------- array backends / configs
@pytest.mark.parametrize("name", ["torch", "numpy"])
def test_array_backend_primitives_agree(name):
    be = get_array_backend(name)
    assert be.name == name
    rows = [[3.0, -1.0, 2.0], [0.0, 5.0, -4.0], [1.0, 1.0, 1.0], [9.0, 0.0, 0.5]]
    X = be.stack([be.asarray(r) for r in rows])
    A = np.asarray(rows)
    f = lambda t: np.asarray(t, dtype=np.float64)
    assert np.allclose(f(be.median(X, axis=0)), np.sort(A, 0)[1])   # lower median on both backends
    assert np.allclose(f(be.mean(X, axis=0)), A.mean(0)) and np.isclose(float(be.mean(X)), A.mean())
    assert np.allclose(f(be.sum(X, axis=1)), A.sum(1)) and np.isclose(float(be.sum(X)), A.sum())
    assert np.allclose(f(be.sort(X, axis=0)), np.sort(A, 0))
    assert np.array_equal(np.asarray(be.argsort(be.asarray([3.0, 1.0, 2.0]))), [1, 2, 0])
    assert np.allclose(f(be.sqrt(be.abs(X))), np.sqrt(np.abs(A)))
    assert np.allclose(f(be.maximum(X, be.asarray(1.0, like=X))), np.maximum(A, 1.0))
    assert np.allclose(f(be.minimum(X, be.asarray(1.0, like=X))), np.minimum(A, 1.0))
    assert tuple(be.reshape(X, (2, 6)).shape) == (2, 6)
    c = be.copy(X)
    c[0, 0] = 100.0
    assert float(X[0, 0]) == 3.0
    assert np.allclose(f(be.matmul(X, be.reshape(X, (3, 4)))), A @ A.reshape(3, 4))
    assert np.allclose(f(be.index_select(X, 0, [2, 0])), A[[2, 0]])
    assert float(be.max(X)) == 9.0


def test_asarray_like_follows_dtype():
    t = get_array_backend("torch").asarray([1, 2], like=torch.zeros(1, dtype=torch.float64))
    assert t.dtype == torch.float64
    a = get_array_backend("numpy").asarray([1, 2], like=np.zeros(1, dtype=np.float32))
    assert a.dtype == np.float32


def test_backend_selection_is_honoured_and_restored():
    assert get_backend().name == "torch"
    with use_backend("numpy"):
        assert get_backend().name == "numpy"
        with use_backend("torch"):
            assert get_backend().name == "torch"
        assert get_backend().name == "numpy"
    assert get_backend().name == "torch"
    with pytest.raises(ValueError, match="unknown backend"):
        set_backend("jax")
    with pytest.raises(RuntimeError):
        with use_backend("numpy"):
            raise RuntimeError("boom")
    assert get_backend().name == "torch"
    set_backend("numpy")
    try:
        assert get_backend().name == "numpy"
    finally:
        set_backend("torch")


def test_dependency_report(monkeypatch):
    monkeypatch.setenv("BYZPY_FORCE_CPU", "1")
    assert _dependencies.preferred_device() == "cpu"
    monkeypatch.delenv("BYZPY_FORCE_CPU")
    monkeypatch.delenv("BYZPY_FORCE_GPU", raising=False)
    assert _dependencies.preferred_device() == ("cuda" if torch.cuda.is_available() else "cpu")
    if not torch.cuda.is_available():
        monkeypatch.setenv("BYZPY_FORCE_GPU", "yes")
        with pytest.raises(RuntimeError, match="BYZPY_FORCE_GPU"):
            _dependencies.preferred_device()
    reqs = _dependencies.base_requirements()
    assert any(r.startswith("torch") for r in reqs) and any(r.startswith("cloudpickle") for r in reqs)


# ------------------------------------------------------------------------------------ legacy transports
def test_local_transport_delivery_counters_and_unregister():
    t = LocalTransport()
    got = []
    t.register("a", got.append)
    t.register("b", lambda m: got.append(("b", m)))
    t.send("a", 1)
    t.send("b", 2)
    t.send("a", 3)
    assert got == [1, ("b", 2), 3] and t.delivered == {"a": 2, "b": 1} and set(t.known_nodes()) == {"a", "b"}
    t.unregister("a")
    t.unregister("a")
    with pytest.raises(KeyError, match="Unknown node_id a"):
        t.send("a", 4)
    with pytest.raises(TypeError):
        t.register("c", "not callable")


def test_local_transport_failed_delivery_is_not_counted():
    t = LocalTransport()

    def bad(_):
        raise RuntimeError("handler failed")

    t.register("x", bad)
    with pytest.raises(RuntimeError):
        t.send("x", 1)
    assert t.delivered["x"] == 0


def test_tcp_mailbox_roundtrip_large_payload_and_timeout():
    box = TcpMailbox()
    try:
        assert box.port > 0
        with pytest.raises(queue.Empty):
            box.recv(timeout=0.05)
        send_message((box.host, box.port), {"k": [1, 2, 3]})
        assert box.recv(timeout=2.0) == {"k": [1, 2, 3]}
        big = torch.arange(300_000, dtype=torch.float32)
        send_message((box.host, box.port), big)
        assert torch.equal(box.recv(timeout=5.0), big)
    finally:
        box.close()


def test_tcp_mailbox_drops_malformed_frames_and_keeps_serving():
    box = TcpMailbox()
    try:
        with socket.create_connection((box.host, box.port)) as s:
            s.sendall(struct.pack(">Q", 10) + b"short")             # truncated body
        with socket.create_connection((box.host, box.port)) as s:
            s.sendall(struct.pack(">Q", 4) + b"\x00\x01\x02\x03")   # not a pickle
        send_message((box.host, box.port), "still alive")
        assert box.recv(timeout=2.0) == "still alive"
    finally:
        box.close()


def test_tcp_transport_routes_by_node_id():
    t = TcpTransport()
    got = {"a": [], "b": []}
    try:
        t.register("a", got["a"].append)
        t.register("b", got["b"].append)
        with pytest.raises(ValueError, match="already registered"):
            t.register("a", print)
        host, port = t.address_of("b")
        assert host == "127.0.0.1" and port != t.address_of("a")[1]
        for i in range(3):
            t.send("a", i)
        t.send("b", "x")
        with pytest.raises(KeyError):
            t.send("zz", 0)
        deadline = time.time() + 5
        while (len(got["a"]) < 3 or not got["b"]) and time.time() < deadline:
            time.sleep(0.01)
        assert sorted(got["a"]) == [0, 1, 2] and got["b"] == ["x"]
    finally:
        t.close()


def test_tcp_transport_survives_handler_exceptions():
    t = TcpTransport()
    seen = []

    def handler(m):
        if m == "bad":
            raise ValueError("nope")
        seen.append(m)

    try:
        t.register("n", handler)
        t.send("n", "bad")
        t.send("n", "good")
        deadline = time.time() + 5
        while not seen and time.time() < deadline:
            time.sleep(0.01)
        assert seen == ["good"]
    finally:
        t.close()


# ---------------------------------------------------------------------------- node runner / cluster
def _count_step(state):
    state["steps"] = state.get("steps", 0) + 1
    return state


def _collect(state, msg):
    state.setdefault("inbox", []).append(msg)
    return state


def test_node_runner_commands_inbox_and_auto_stepping():
    r = NodeRunner(_count_step, _collect, init_state={"steps": 10})
    r.start()
    try:
        assert r.state() == {"steps": 10}
        r.step()
        r.step()
        assert r.state()["steps"] == 12
        r.send_message({"g": torch.ones(2)})
        r.send_message("second")
        st = r.state()                                              # inbox is drained before every command
        assert st["inbox"][1] == "second" and torch.equal(st["inbox"][0]["g"], torch.ones(2))
        r.start_auto(0.01)
        time.sleep(0.25)
        r.stop_auto()
        n = r.state()["steps"]
        assert n > 12                                               # at least one automatic step happened
        time.sleep(0.05)
        assert r.state()["steps"] == n                              # auto-stepping really stopped
        assert r._command("bogus")[0] == "error"
    finally:
        r.stop()
    assert not r._proc.is_alive()


def test_node_runner_background_pump_thread():
    r = NodeRunner(_count_step, _collect)
    r.start()
    try:
        r.start_async(0.005)
        r.start_async(0.005)                                        # second call is a no-op
        time.sleep(0.2)
        r._stop_pump_thread()
        assert r.state()["steps"] >= 3 and r._pump_thread is None
    finally:
        r.stop()


@pytest.mark.parametrize("transport", [None, "local"])
def test_node_cluster_membership_messaging_and_state(transport):
    tr = LocalTransport() if transport else None
    c = NodeCluster(transport=tr)
    c.add_node("a", _count_step, _collect)
    c.add_node("b", _count_step, _collect, init_state={"steps": 5})
    with pytest.raises(ValueError, match="already exists"):
        c.add_node("a", _count_step, _collect)
    assert len(c) == 2 and list(c) == ["a", "b"]
    c.start_all()
    try:
        c.send("a", "hello")
        c.send("b", 42)
        assert c.state("a")["inbox"] == ["hello"] and c.state("b") == {"steps": 5, "inbox": [42]}
        with pytest.raises(KeyError):
            c.state("zz") if tr is None else c.send("zz", 1)
        c.start_auto("b", 0.01)
        c.barrier(0.15)
        assert c.state("b")["steps"] > 5
        if tr is not None:
            assert tr.delivered == {"a": 1, "b": 1}
    finally:
        c.stop_all()


def test_parameter_server_runner_custom_aggregator_and_facade():
    from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian

    fns = [lambda k=k: torch.full((3,), float(k)) for k in (1, 2, 30)]
    runner = ParameterServerRunner(fns, aggregator=CoordinateWiseMedian().aggregate)
    assert runner.worker_ids == ["w0", "w1", "w2"] and runner.server_id == "server" and len(runner.cluster) == 4
    runner.start()
    try:
        assert torch.equal(runner.run_round(), torch.full((3,), 2.0))
        assert torch.equal(runner.run_round(), torch.full((3,), 2.0))    # inbox was reset between rounds
    finally:
        runner.stop()
    assert torch.equal(mean_aggregate([torch.ones(2), 3 * torch.ones(2)]), torch.full((2,), 2.0))


def test_decentralized_parameter_server_facade():
    class N:
        def __init__(self, v):
            self.grad = torch.full((2,), float(v))

    async def go():
        ps = DecentralizedParameterServer([N(1), N(3)], [N(100)], mean_aggregate)
        assert isinstance(ps.runner, ParameterServerRunner) and ps.rounds == 0
        await ps.bootstrap()
        try:
            out = await ps.round()
            out2 = await ps.round()
        finally:
            await ps.shutdown()
        return out, out2, ps.rounds

    out, out2, rounds = run(go())
    assert torch.equal(out, torch.full((2,), 2.0)) and torch.equal(out2, out) and rounds == 2


# --------------------------------------------------------------------------------------------- tracing
def test_tracer_tags_summary_and_exceptions():
    tr = Tracer(cuda=False)
    with tr.span("a", op="x"):
        time.sleep(0.002)
    with tr.span("a"):
        pass
    with pytest.raises(KeyError):
        with tr.span("b", kind="err"):
            raise KeyError("inside")
    recs = tr.finalize()
    assert [r["name"] for r in recs] == ["a", "a", "b"] and recs[0]["op"] == "x" and recs[0]["host_ms"] >= 1.0
    s = tr.summary()
    assert s["a"]["calls"] == 2 and s["b"]["calls"] == 1 and s["a"]["host_ms"] >= recs[0]["host_ms"]
    assert "device_ms" not in recs[0]
    with nvtx_range("noop"):
        pass


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_cuda_time_requires_cuda():
    with pytest.raises(Exception):
        cuda_time_ms(lambda: None)


# ---------------------------------------------------------------------------------------- import alias
def test_alias_install_uninstall_cycle():
    import byzpy_b200.compat as compat

    had = [m for m in sys.modules if m == "byzpy" or m.startswith("byzpy.")]
    if had:
        pytest.skip("a real byzpy is imported in this process")
    compat.install_alias()
    compat.install_alias()                                          # idempotent
    try:
        import byzpy.aggregators.coordinate_wise as cw
        from byzpy.engine.graph.pool import ActorPoolConfig

        import byzpy_b200.aggregators.coordinate_wise as real
        from byzpy_b200.engine.graph.pool import ActorPoolConfig as RealCfg

        assert cw is real and ActorPoolConfig is RealCfg
        with pytest.raises(ImportError):
            import byzpy.does_not_exist  # noqa: F401
    finally:
        compat.uninstall_alias()
        compat.uninstall_alias()
    assert not [m for m in sys.modules if m == "byzpy" or m.startswith("byzpy.")]
    with pytest.raises(ImportError):
        import byzpy  # noqa: F401


# ---------------------------------------------------------------------------- attack subtask plumbing
class _InlinePool:
    size = 4

    async def run_subtask(self, st):
        return st.run()

    def worker_affinities(self):
        return ()


@pytest.mark.parametrize("mk,inputs", [
    (lambda: SignFlipAttack(scale=-3.0, chunk_size=7), lambda vs: {"base_grad": vs[0]}),
    (lambda: EmpireAttack(scale=-1.0, chunk_size=2), lambda vs: {"honest_grads": vs}),   # (gradients per subtask)
    (lambda: LittleAttack(f=1, chunk_size=7), lambda vs: {"honest_grads": vs}),
    (lambda: InfAttack(chunk_size=7), lambda vs: {"honest_grads": vs}),
    (lambda: MimicAttack(epsilon=1, chunk_size=7), lambda vs: {"honest_grads": vs}),
    (lambda: GaussianAttack(mu=1.0, sigma=0.0, chunk_size=7), lambda vs: {"honest_grads": vs}),
])
def test_attack_subtask_path_equals_direct(mk, inputs):
    g = torch.Generator().manual_seed(0)
    vs = [torch.randn(200, generator=g) for _ in range(5)]
    ctx = OpContext("n", metadata={"pool_size": 4})
    atk = mk()
    direct = atk.compute(inputs(vs), context=ctx)
    sts = list(atk.create_subtasks(inputs(vs), context=ctx)) if atk.supports_subtasks else []
    pooled = run(atk.run(inputs(vs), context=ctx, pool=_InlinePool()))
    assert pooled.shape == direct.shape and torch.allclose(pooled, direct, equal_nan=True)
    if atk.supports_subtasks:
        assert len(sts) > 1


def test_sign_flip_subtasks_without_base_grad_fall_back_and_fold():
    atk = SignFlipAttack(scale=-2.0)
    assert list(atk.create_subtasks({}, context=OpContext("n"))) == []
    v = torch.arange(6.0).reshape(2, 3)
    out = atk.reduce_subtasks([], {"base_grad": v}, context=OpContext("n"))
    assert torch.equal(out, -2.0 * v)
    fold = atk.fold(4)
    assert fold.kind == "scale" and fold.scale == -2.0


# ------------------------------------------------------------------------------------------------- CLI
def test_cli_human_outputs_and_parser(capsys):
    assert cli_main(["doctor"]) == 0
    out = capsys.readouterr().out
    assert "python_version:" in out and "torch:" in out and "  - available: True" in out
    assert cli_main(["list", "pre-aggregators", "--format", "json"]) == 0
    items = json.loads(capsys.readouterr().out)
    assert items["component"] == "pre-aggregators"
    assert "byzpy_b200.pre_aggregators.bucketing.Bucketing" in items["items"]           # qualified, like the reference
    assert {n.rsplit(".", 1)[-1] for n in items["items"]} == {"ARC", "Bucketing", "Clipping", "NearestNeighborMixing"}
    assert cli_main(["list", "attacks", "--format", "json", "--short"]) == 0
    assert {"EmpireAttack", "GaussianAttack", "InfAttack", "LabelFlipAttack", "LittleAttack", "MimicAttack",
            "SignFlipAttack"} <= set(json.loads(capsys.readouterr().out)["items"])
    with pytest.raises(SystemExit):
        cli_main([])
    with pytest.raises(SystemExit):
        cli_main(["list", "optimizers"])
    ns = build_parser().parse_args(["bench", "--", "--steps", "2"])
    assert ns.rest == ["--", "--steps", "2"]


def test_cli_subclass_discovery_skips_abstract_and_private():
    from byzpy_b200.aggregators.base import Aggregator

    qualified = _load_subclasses("byzpy_b200.aggregators", Aggregator)
    assert "byzpy_b200.aggregators.coordinate_wise.median.CoordinateWiseMedian" in qualified
    assert not any(".tests" in q for q in qualified)
    names = [q.rsplit(".", 1)[-1] for q in qualified]
    assert qualified == sorted(qualified) and "Aggregator" not in names and "GramAggregator" not in names
    assert len(names) >= 12


def test_ndarray_backend_module_paths_and_protocol(byzpy_alias):
    """The reference's import paths ``engine.backend.ndarray.{base,torch}`` (reference base.py:8-27,
    torch.py:10-71) resolve, directly and through the ``byzpy`` alias, and both implementations satisfy the
    16-primitive protocol."""
    import importlib

    import numpy as np

    nd_base = importlib.import_module("byzpy_b200.engine.backend.ndarray.base")
    nd_numpy = importlib.import_module("byzpy_b200.engine.backend.ndarray.numpy")
    nd_torch = importlib.import_module("byzpy_b200.engine.backend.ndarray.torch")
    assert importlib.import_module("byzpy.engine.backend.ndarray.base")._Backend is nd_base._Backend
    assert importlib.import_module("byzpy.engine.backend.ndarray.torch")._TorchBackend is nd_torch._TorchBackend
    prims = [n for n in vars(nd_base._Backend) if not n.startswith("_") and callable(getattr(nd_base._Backend, n))]
    assert len(prims) == 16
    for impl, make in ((nd_torch._TorchBackend(), lambda v: torch.tensor(v)), (nd_numpy._NumpyBackend(), np.asarray)):
        assert isinstance(impl, nd_base._Backend) and all(callable(getattr(impl, p)) for p in prims)
        x = impl.stack([make([3.0, 1.0, 2.0]), make([0.0, 5.0, 4.0])], axis=0)
        assert [float(v) for v in impl.median(x, axis=0)] == [0.0, 1.0, 2.0]                 # lower median of two rows
        assert [float(v) for v in impl.sum(impl.maximum(x, impl.copy(x) * 0), axis=1)] == [6.0, 9.0]
        assert [int(v) for v in impl.argsort(make([3.0, 1.0, 2.0]))] == [1, 2, 0]
    # importing the submodule called ``torch`` rebinds the package attribute of that name; the package's own code
    # must keep working afterwards (it refers to the library through a private alias)
    from byzpy_b200.engine.backend.ndarray import get_array_backend

    be = get_array_backend("torch")
    assert be.stack([torch.ones(2), torch.zeros(2)]).shape == (2, 2)


def test_cli_build_and_bench_subcommands(capsys):
    """``byzpy-b200 build`` returns the in-tree extension (already built: nothing recompiles); ``byzpy-b200 bench``
    forwards its arguments to bench.py, which on a box without a GPU says so and exits 1."""
    from byzpy_b200 import cli

    pytest.importorskip("byzpy_b200._C")
    assert cli.main(["build"]) == 0
    assert "_C" in capsys.readouterr().out
    if not torch.cuda.is_available():            # bench.py refuses politely without a device (exit code 1)
        assert cli.main(["bench", "--", "--steps", "1", "--warmup", "0"]) == 1
    assert cli.main(["list", "attacks"]) == 0
    out = capsys.readouterr().out
    assert "byzpy_b200.attacks" in out and "SignFlipAttack" in out
