"""Orchestrators: ParameterServer (generic actor path, every aggregator, with and without a
pool -- the reference fails here for Median/TrimmedMean/GeometricMedian/CenteredClipping),
PeerToPeer, legacy runners/transports, CLI, utils, shared store, arenas."""
import asyncio
import json

import pytest
import torch
import torch.nn as nn

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean
from byzpy_b200.aggregators.geometric_wise import GeometricMedian, MultiKrum
from byzpy_b200.aggregators.norm_wise import CenteredClipping
from byzpy_b200.attacks import EmpireAttack, SignFlipAttack
from byzpy_b200.cli import main as cli_main
from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig
from byzpy_b200.engine.node.actors import ByzantineNodeActor, HonestNodeActor
from byzpy_b200.engine.node.base import ByzantineNode, HonestNode
from byzpy_b200.engine.node.device import DeviceByzantineNode, DeviceHonestNode
from byzpy_b200.engine.node.mixin import P2PByzantineMixin, P2PHonestMixin
from byzpy_b200.engine.parameter_server.ps import ParameterServer
from byzpy_b200.engine.parameter_server.runner import ParameterServerRunner
from byzpy_b200.engine.peer_to_peer.topology import Topology
from byzpy_b200.engine.peer_to_peer.train import PeerToPeer
from byzpy_b200.engine.storage.shared_store import (SharedTensorHandle, cleanup_tensor, materialize,
                                                    open_tensor, register_tensor)
from byzpy_b200.engine.transport import LocalTransport, TcpTransport
from byzpy_b200.parallel.arena import ParamArena, flatten_grads, flatten_params
from byzpy_b200.pre_aggregators import Bucketing
from byzpy_b200.utils import train_with_progress


def run(coro):
    return asyncio.run(coro)


class Hon(HonestNode):
    def __init__(self, seed, d=12):
        self.g = torch.Generator().manual_seed(seed)
        self.d = d
        self.applied = []

    def next_batch(self):
        return torch.randn(self.d, generator=self.g), torch.zeros(1)

    def honest_gradient(self, x, y):
        return x + 1.0

    def apply_server_gradient(self, g):
        self.applied.append(g.clone())

    def history(self):
        return self.applied


class Byz(ByzantineNode):
    def __init__(self):
        self.attack = EmpireAttack(scale=-3.0)
        self.applied = []

    def next_batch(self):
        return torch.empty(0), torch.empty(0)

    def byzantine_gradient(self, x, y, honest_grads=None):
        return self.attack.apply(honest_grads=list(honest_grads))

    def apply_server_gradient(self, g):
        self.applied.append(g)

    def history(self):
        return self.applied


AGGS = [lambda: CoordinateWiseMedian(), lambda: CoordinateWiseTrimmedMean(f=1), lambda: GeometricMedian(),
        lambda: CenteredClipping(c_tau=1.0), lambda: MultiKrum(f=1, q=2)]


@pytest.mark.parametrize("mk", AGGS)
@pytest.mark.parametrize("use_pool", [False, True])
def test_parameter_server_round_generic_path(mk, use_pool):
    async def scenario():
        hon = [await HonestNodeActor.spawn(Hon, backend="thread", args=(i,)) for i in range(4)]
        byz = [await ByzantineNodeActor.spawn(Byz, backend="thread")]
        pool = None
        if use_pool:
            pool = ActorPool([ActorPoolConfig(backend="thread", count=2)])
            await pool.start()
        ps = ParameterServer(hon, byz, mk(), update_byzantines=False, actor_pool=pool)
        g = await ps.round()
        # deterministic (submission-order) gather: recompute what the PS saw
        mirrors = [Hon(i) for i in range(4)]
        rows = [m.honest_gradient_for_next_batch() for m in mirrors]
        rows.append(Byz().byzantine_gradient_for_next_batch(rows))
        assert torch.allclose(g, mk().aggregate(rows), rtol=1e-5, atol=1e-6)
        for h in hon:
            assert len(await h.history()) == 1
        assert len(await byz[0].history()) == 0
        await ps.shutdown()
        if pool is not None:
            await pool.shutdown()

    run(scenario())


def test_parameter_server_pre_aggregator_and_update_byzantines_plain_objects():
    hon, byz = [Hon(i) for i in range(4)], [Byz()]
    ps = ParameterServer(hon, byz, CoordinateWiseMedian(), pre_aggregator=Bucketing(2, perm=[0, 1, 2, 3, 4]),
                         update_byzantines=True)
    g = ps.round_sync()
    assert g.shape == (12,) and len(byz[0].applied) == 1 and ps.rounds == 1
    assert ps.device_round is None
    with pytest.raises(RuntimeError):
        ps.step()


def test_device_nodes_fall_back_to_generic_round_on_cpu():
    torch.manual_seed(0)

    def src():
        return torch.randn(8, 6), torch.randint(0, 3, (8,))

    hon = [DeviceHonestNode(nn.Linear(6, 3), data=src, lr=0.1, momentum=0.0, device="cpu") for _ in range(3)]
    byz = [DeviceByzantineNode(SignFlipAttack(), model=nn.Linear(6, 3), data=src, device="cpu")]
    ps = ParameterServer(hon, byz, CoordinateWiseMedian(), update_byzantines=True)
    assert ps.device_round is None
    before = flatten_params(hon[0].model).clone()
    g = ps.round_sync()
    after = flatten_params(hon[0].model)
    assert torch.allclose(after, before - 0.1 * g, atol=1e-6)
    with pytest.raises(RuntimeError):
        ParameterServer(hon, byz, CoordinateWiseMedian(), fused=True)
    sd = hon[0].dump_state_dict()
    nn.Linear(6, 3).load_state_dict(sd, strict=True)


class PH(P2PHonestMixin):
    def __init__(self, seed):
        torch.manual_seed(0)
        self.model = nn.Linear(5, 2)
        self.device = torch.device("cpu")
        self.criterion = nn.CrossEntropyLoss()
        self.p2p_agg = CoordinateWiseTrimmedMean(f=1)
        self.g = torch.Generator().manual_seed(seed)

    def next_batch(self):
        return torch.randn(16, 5, generator=self.g), torch.randint(0, 2, (16,), generator=self.g)

    def params(self):
        return self.get_param_vector()


class PB(P2PByzantineMixin):
    def __init__(self):
        self.device = torch.device("cpu")
        self.attack = EmpireAttack(scale=-5.0)


def test_peer_to_peer_round_writes_back_robust_aggregate():
    async def scenario():
        hon = [await HonestNodeActor.spawn(PH, backend="thread", args=(i,)) for i in range(4)]
        byz = [await ByzantineNodeActor.spawn(PB, backend="thread")]
        p2p = PeerToPeer(hon, byz, Topology.complete(5), lr=0.1)
        await p2p.bootstrap()
        p0 = await hon[0].params()
        await p2p.round()
        p1 = await hon[0].params()
        # manual expectation for node 0
        mirrors = [PH(i) for i in range(4)]
        halves = [m.p2p_half_step(0.1) for m in mirrors]
        mal = PB().p2p_broadcast_vector(neighbor_vectors=halves, like=halves[0])
        exp = CoordinateWiseTrimmedMean(f=1).aggregate([halves[0]] + halves[1:] + [mal])
        assert not torch.equal(p0, p1) and torch.allclose(p1, exp, atol=1e-6)
        assert p2p.runner.rounds == 1
        await p2p.shutdown()
        for a in hon + byz:
            await a.close()

    run(scenario())


def test_p2p_default_context_follows_the_environment_switch(monkeypatch):
    """``BYZPY_P2P_CONTEXT=process`` gives the reference's default (a ``ProcessContext`` per node) and the round still
    produces the robust aggregate; unset, nodes sit in ``InProcessContext``; anything else is rejected."""
    from byzpy_b200.engine.node.context import InProcessContext, ProcessContext

    async def scenario(expect_cls):
        hon = [await HonestNodeActor.spawn(PH, backend="thread", args=(i,)) for i in range(3)]
        p2p = PeerToPeer(hon, [], Topology.complete(3), lr=0.1)
        await p2p.bootstrap()
        try:
            assert all(isinstance(n.context, expect_cls) for n in p2p._runner._cluster.nodes.values())
            await p2p.round()
            halves = [PH(i).p2p_half_step(0.1) for i in range(3)]
            exp = CoordinateWiseTrimmedMean(f=1).aggregate(halves)
            assert torch.allclose(await hon[0].params(), exp, atol=1e-6)
        finally:
            await p2p.shutdown()
            for a in hon:
                await a.close()

    monkeypatch.delenv("BYZPY_P2P_CONTEXT", raising=False)
    run(scenario(InProcessContext))
    monkeypatch.setenv("BYZPY_P2P_CONTEXT", "process")
    run(scenario(ProcessContext))
    monkeypatch.setenv("BYZPY_P2P_CONTEXT", "bogus")
    with pytest.raises(ValueError):
        run(scenario(InProcessContext))


def test_decentralized_peer_to_peer_round_matches_manual_expectation():
    """``DecentralizedPeerToPeer`` over DecentralizedNodes: the round writes the robust aggregate of
    own + neighbour half-steps back into every honest model (ring topology, message-count driven)."""
    from byzpy_b200.engine.peer_to_peer.runner import DecentralizedPeerToPeer

    async def scenario():
        hon = [await HonestNodeActor.spawn(PH, backend="thread", args=(i,)) for i in range(4)]
        byz = [await ByzantineNodeActor.spawn(PB, backend="thread")]
        p2p = DecentralizedPeerToPeer(hon, byz, Topology.complete(5), lr=0.1, recv_timeout=10.0)
        await p2p.start()
        await p2p.run_round_async()
        got = [await h.params() for h in hon]
        mirrors = [PH(i) for i in range(4)]
        halves = [m.p2p_half_step(0.1) for m in mirrors]
        mal = PB().p2p_broadcast_vector(neighbor_vectors=halves, like=halves[0])
        for i in range(4):
            others = [halves[j] for j in range(4) if j != i] + [mal]
            exp = CoordinateWiseTrimmedMean(f=1).aggregate([halves[i]] + others)
            assert torch.allclose(got[i], exp, atol=1e-6)
        assert p2p.rounds == 1
        await p2p.run_round_async()
        assert p2p.rounds == 2
        await p2p.stop()
        for a in hon + byz:
            await a.close()

    run(scenario())


class _DuckAggregateOnly:
    """An honest gossip node from before the mixins: ``p2p_half_step`` and ``p2p_aggregate(vectors)`` only."""

    def __init__(self, start):
        self.theta = torch.tensor([float(start), float(start)])
        self.seen = None

    def p2p_half_step(self, lr):
        self.theta = self.theta - lr
        return self.theta.clone()

    def p2p_aggregate(self, vectors):
        self.seen = len(vectors)
        self.theta = torch.stack(list(vectors)).median(0).values
        return self.theta


class _DuckBare:
    """A node with neither aggregation hook: the runner aggregates with a coordinate-wise median and loads the
    result through ``set_param_vector``."""

    def __init__(self, start):
        self.theta = torch.tensor([float(start), float(start)])

    def p2p_half_step(self, lr):
        return self.theta.clone()

    def set_param_vector(self, vec):
        self.theta = torch.as_tensor(vec).clone()


class _DuckBroken(_DuckAggregateOnly):
    def p2p_aggregate(self, vectors):
        return self.no_such_attribute          # an AttributeError raised INSIDE the hook must not be mistaken for "no hook"


def test_decentralized_peer_to_peer_accepts_duck_typed_nodes():
    from byzpy_b200.engine.peer_to_peer.runner import DecentralizedPeerToPeer

    async def scenario(nodes):
        p2p = DecentralizedPeerToPeer(nodes, [], Topology.complete(len(nodes)), lr=1.0, recv_timeout=5.0)
        await p2p.start()
        try:
            await p2p.run_round_async()
        finally:
            await p2p.stop()

    old = [_DuckAggregateOnly(s) for s in (1.0, 5.0, 9.0)]
    run(scenario(old))
    assert all(n.seen == 3 for n in old)                                   # own + two neighbours
    assert all(torch.equal(n.theta, torch.tensor([4.0, 4.0])) for n in old)   # median of (0, 4, 8)
    bare = [_DuckBare(s) for s in (1.0, 5.0, 9.0)]
    run(scenario(bare))
    assert all(torch.equal(n.theta, torch.tensor([5.0, 5.0])) for n in bare)
    with pytest.raises(AttributeError):
        run(scenario([_DuckBroken(s) for s in (1.0, 2.0, 3.0)]))


def test_gram_family_subtasks_chunk_the_feature_dimension_coarsely():
    """Regression: the reference's row-chunk default (chunk_size=32) must not be used as a FEATURE
    chunk -- a 1.2 M-parameter gradient became 37 500 subtasks."""
    from byzpy_b200.aggregators.geometric_wise import MultiKrum
    from byzpy_b200.engine.graph.operator import OpContext

    op = MultiKrum(f=1, q=3)
    grads = [torch.randn(300_000) for _ in range(5)]
    ctx = OpContext(node_name="agg", metadata={"pool_size": 2})
    tasks = list(op.create_subtasks({"gradients": grads}, context=ctx))
    assert 1 <= len(tasks) <= 64
    op._release_packed(grads) if hasattr(op, "_release_packed") else None


def g1():
    return torch.tensor([1.0, 2.0])


def g2():
    return torch.tensor([3.0, 4.0])


@pytest.mark.parametrize("transport", [None, LocalTransport, TcpTransport])
def test_parameter_server_runner_mean_round(transport):
    tr = transport() if transport else None
    r = ParameterServerRunner([g1, g2], transport=tr)
    r.start()
    try:
        assert torch.equal(r.run_round(), torch.tensor([2.0, 3.0]))
        assert torch.equal(r.run_round(), torch.tensor([2.0, 3.0]))
    finally:
        r.stop()
        if hasattr(tr, "close"):
            tr.close()


def test_local_transport_unknown_node():
    with pytest.raises(KeyError):
        LocalTransport().send("nobody", 1)


def test_cli(capsys):
    assert cli_main(["version"]) == 0
    assert capsys.readouterr().out.strip()
    assert cli_main(["doctor", "--format", "json"]) == 0
    rep = json.loads(capsys.readouterr().out)
    assert rep["torch"]["available"] and "kernels" in rep
    assert cli_main(["list", "aggregators", "--format", "json"]) == 0
    items = [q.rsplit(".", 1)[-1] for q in json.loads(capsys.readouterr().out)["items"]]
    assert {"CoordinateWiseMedian", "MultiKrum", "Krum", "CAF", "SMEA", "MoNNA"} <= set(items)
    assert "CoordinateWiseAggregator" not in items
    assert cli_main(["list", "attacks"]) == 0 and "LittleAttack" in capsys.readouterr().out
    assert cli_main(["list", "pre-aggregators"]) == 0 and "Bucketing" in capsys.readouterr().out


def test_train_with_progress():
    class FakePS:
        n = 0

        async def round(self):
            self.n += 1

    evals = []
    ps = FakePS()
    hist = run(train_with_progress(ps, 10, eval_callback=lambda: evals.append(ps.n) or {"acc": 0.5}, eval_interval=5))
    assert ps.n == 10 and evals == [5, 10]
    assert hist == [{"round": 5, "metrics": {"acc": 0.5}}, {"round": 10, "metrics": {"acc": 0.5}}]

    async def scored(round_num):                                   # the reference's callback shape: async, takes the round
        return {"acc": round_num / 10}

    hist = run(train_with_progress(FakePS(), 4, eval_callback=scored, eval_interval=2))
    assert [h["round"] for h in hist] == [2, 4] and hist[1]["metrics"] == {"acc": 0.4}
    assert run(train_with_progress(FakePS(), 3)) == []


def test_shared_store_roundtrip():
    x = torch.arange(12.0).reshape(3, 4)
    h = register_tensor(x)
    assert isinstance(h, SharedTensorHandle) and h.shape == (3, 4) and h.dtype == "float32"
    with open_tensor(h) as arr:
        assert arr[2, 3] == 11.0
    assert torch.equal(materialize({"name": h.name, "shape": h.shape, "dtype": h.dtype}), x)
    cleanup_tensor(h)
    cleanup_tensor(h)  # idempotent


def test_param_arena_views_and_layout():
    m = nn.Sequential(nn.Linear(3, 4), nn.Linear(4, 2))
    ref = flatten_params(m).clone()
    arena = ParamArena(m)
    assert arena.d == ref.numel() and arena.d_pad % 1024 == 0 and arena.check_bound()
    assert torch.equal(arena.param_vector(), ref)
    m(torch.randn(5, 3)).sum().backward()
    assert torch.equal(arena.grad_vector(), flatten_grads(m)) and arena.check_bound()
    arena.flat_params[:3] = 7.0
    assert torch.equal(next(m.parameters()).reshape(-1)[:3], torch.full((3,), 7.0))
    arena.zero_grad()
    assert all(p.grad.abs().sum() == 0 for p in m.parameters())


class Flaky(Hon):
    def __init__(self, seed, mode):
        super().__init__(seed)
        self.mode = mode

    def honest_gradient(self, x, y):
        if self.mode == "raise":
            raise RuntimeError("node crashed")
        if self.mode == "hang":
            import time

            time.sleep(1.0)
        return super().honest_gradient(x, y)


def test_parameter_server_failure_detection_and_fault_injection():
    async def scenario():
        hon = [await HonestNodeActor.spawn(Hon, backend="thread", args=(i,)) for i in range(3)]
        hon.append(await HonestNodeActor.spawn(Flaky, backend="thread", args=(7, "raise")))
        hon.append(await HonestNodeActor.spawn(Flaky, backend="thread", args=(8, "hang")))
        strict = ParameterServer(hon[:4], [], CoordinateWiseMedian())
        with pytest.raises(Exception):
            await strict.round()
        ps = ParameterServer(hon, [], CoordinateWiseMedian(), node_timeout=0.2, tolerate_failures=True)
        g = await ps.round()
        rows = [Hon(i).honest_gradient_for_next_batch() for i in range(3)]
        # the strict server consumed one batch of nodes 0..2 already: regenerate their second batch
        mirrors = [Hon(i) for i in range(3)]
        for m in mirrors:
            m.honest_gradient_for_next_batch()
        rows = [m.honest_gradient_for_next_batch() for m in mirrors]
        assert torch.allclose(g, CoordinateWiseMedian().aggregate(rows))
        kinds = sorted(f[1] for f in ps.failed if f[1].startswith("honest"))
        assert kinds == ["honest:3", "honest:4"]
        await asyncio.sleep(1.0)
        for a in hon:
            await a.close()

    run(scenario())


def test_checkpoint_roundtrip_generic_nodes(tmp_path):
    from byzpy_b200.utils.checkpoint import load_checkpoint, save_checkpoint

    torch.manual_seed(0)

    def src():
        return torch.randn(8, 6), torch.randint(0, 3, (8,))

    def build():
        hon = [DeviceHonestNode(nn.Linear(6, 3), data=src, lr=0.1, momentum=0.9, device="cpu") for _ in range(3)]
        return ParameterServer(hon, [], CoordinateWiseMedian())

    ps = build()
    for _ in range(3):
        ps.round_sync()
    path = str(tmp_path / "ck.pt")
    save_checkpoint(path, ps)
    ps2 = build()
    assert load_checkpoint(path, ps2) == 3 and ps2.rounds == 3
    for a, b in zip(ps.hon, ps2.hon):
        assert torch.equal(flatten_params(a.model), flatten_params(b.model))
    blob = torch.load(path, weights_only=False)
    # the per-node snapshot is exactly the reference's dump_state_dict() mapping
    nn.Linear(6, 3).load_state_dict(blob["nodes"][0]["state_dict"], strict=True)


def test_checkpoint_resume_continues_the_same_trajectory_and_rejects_mismatches(tmp_path):
    """Training 2 rounds, checkpoint, 2 more rounds == restoring the checkpoint into fresh nodes and training 2 rounds:
    momentum buffers and the RNG stream are part of the checkpoint.  Files of another format or with a different
    number of nodes are refused."""
    from byzpy_b200.utils.checkpoint import FORMAT, load_checkpoint, load_reference_state_dict, save_checkpoint

    def build(k=3):
        torch.manual_seed(0)
        hon = [DeviceHonestNode(nn.Linear(6, 3), data=lambda: (torch.randn(8, 6), torch.randint(0, 3, (8,))), lr=0.1,
                                momentum=0.9, device="cpu") for _ in range(k)]
        return ParameterServer(hon, [], CoordinateWiseMedian())

    ps = build()
    for _ in range(2):
        ps.round_sync()
    path = str(tmp_path / "ck.pt")
    save_checkpoint(path, ps, extra={"note": "after two rounds"})
    for _ in range(2):
        ps.round_sync()
    want = [flatten_params(n.model).clone() for n in ps.hon]

    resumed = build()
    torch.manual_seed(12345)                           # a different stream until the checkpoint restores the saved one
    assert load_checkpoint(path, resumed) == 2
    for _ in range(2):
        resumed.round_sync()
    assert resumed.rounds == 4
    for a, b in zip(want, [flatten_params(n.model) for n in resumed.hon]):
        assert torch.allclose(a, b, atol=1e-6)

    assert torch.load(path, weights_only=True)["extra"] == {"note": "after two rounds"}
    with pytest.raises(ValueError):
        load_checkpoint(path, build(k=2))                                   # fewer nodes than the file
    bad = str(tmp_path / "bad.pt")
    torch.save({"format": "something-else", "nodes": []}, bad)
    with pytest.raises(ValueError):
        load_checkpoint(bad, build())
    assert FORMAT.startswith("byzpy_b200.ckpt")
    fresh = nn.Linear(6, 3)
    load_reference_state_dict(fresh, ps.hon[0].dump_state_dict())
    assert torch.equal(flatten_params(fresh), flatten_params(ps.hon[0].model))


def test_tracer_records_graph_nodes():
    from byzpy_b200.engine.graph.ops import make_single_operator_graph
    from byzpy_b200.engine.graph.scheduler import NodeScheduler
    from byzpy_b200.utils.tracing import Tracer, nvtx_range

    tr = Tracer(cuda=False)
    g = make_single_operator_graph(node_name="agg", operator=CoordinateWiseMedian(), input_keys=("gradients",))
    run(NodeScheduler(g, metadata={"tracer": tr}).run({"gradients": [torch.randn(5) for _ in range(3)]}))
    summ = tr.summary()
    assert summ["node:agg"]["calls"] == 1 and summ["node:agg"]["host_ms"] > 0
    with nvtx_range("noop"):
        pass


def test_byzpy_import_alias_maps_reference_paths():
    import sys

    from byzpy_b200 import compat

    saved = {k: v for k, v in sys.modules.items() if k == "byzpy" or k.startswith("byzpy.")}
    for k in saved:
        del sys.modules[k]
    compat.install_alias()
    try:
        from byzpy.aggregators.coordinate_wise import CoordinateWiseMedian as A  # type: ignore
        from byzpy.engine.graph.pool import ActorPoolConfig as P  # type: ignore
        from byzpy.engine.peer_to_peer.topology import Topology as T  # type: ignore
        import byzpy  # type: ignore

        from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
        from byzpy_b200.engine.graph.pool import ActorPoolConfig
        from byzpy_b200.engine.peer_to_peer.topology import Topology

        assert A is CoordinateWiseMedian and P is ActorPoolConfig and T is Topology
        assert hasattr(byzpy, "run_operator")
    finally:
        compat.uninstall_alias()
        sys.modules.update(saved)


def test_public_api_surface_of_the_reference_is_importable():
    import importlib

    surface = {
        "byzpy_b200": ["run_operator", "OperatorExecutor", "__version__"],
        "byzpy_b200.aggregators": ["Aggregator"],
        "byzpy_b200.aggregators.coordinate_wise": ["MeanOfMedians", "CoordinateWiseMedian", "CoordinateWiseTrimmedMean"],
        "byzpy_b200.aggregators.geometric_wise": ["GeometricMedian", "Krum", "MultiKrum", "MinimumDiameterAveraging", "MoNNA", "SMEA"],
        "byzpy_b200.aggregators.norm_wise": ["CAF", "CenteredClipping", "ComparativeGradientElimination"],
        "byzpy_b200.aggregators._chunking": ["select_adaptive_chunk_size"],
        "byzpy_b200.aggregators.coordinate_wise._tiling": ["flatten_gradients"],
        "byzpy_b200.pre_aggregators": ["PreAggregator", "Bucketing", "NearestNeighborMixing", "Clipping", "ARC"],
        "byzpy_b200.attacks": ["Attack", "EmpireAttack", "LittleAttack", "SignFlipAttack", "LabelFlipAttack", "GaussianAttack", "InfAttack", "MimicAttack"],
        "byzpy_b200.configs.actor": ["set_actor"],
        "byzpy_b200.configs.backend": ["set_backend", "get_backend", "use_backend"],
        "byzpy_b200.engine": ["CallableOp", "RemoteCallableOp", "make_single_operator_graph", "NodeCluster", "NodeRunner"],
        "byzpy_b200.engine.graph.graph": ["ComputationGraph", "GraphInput", "GraphNode", "graph_input"],
        "byzpy_b200.engine.graph.operator": ["OpContext", "Operator", "MessageTriggerOp"],
        "byzpy_b200.engine.graph.subtask": ["SubTask"],
        "byzpy_b200.engine.graph.scheduler": ["NodeScheduler", "MessageAwareNodeScheduler", "MessageSource"],
        "byzpy_b200.engine.graph.parallel_scheduler": ["ParallelScheduler"],
        "byzpy_b200.engine.graph.pool": ["ActorPool", "ActorPoolConfig", "ActorPoolChannel"],
        "byzpy_b200.engine.graph.ops": ["CallableOp", "RemoteCallableOp", "make_single_operator_graph"],
        "byzpy_b200.engine.graph.lazy": ["GraphBuilder", "LazyNode"],
        "byzpy_b200.engine.graph.session": ["ExecutionSession", "ExecutionFuture"],
        "byzpy_b200.engine.graph.executor": ["OperatorExecutor", "run_operator"],
        "byzpy_b200.engine.actor.base": ["ActorBackend", "ActorRef"],
        "byzpy_b200.engine.actor.channels": ["Endpoint", "ChannelRef", "open_channel"],
        "byzpy_b200.engine.actor.factory": ["resolve_backend"],
        "byzpy_b200.engine.actor.router": ["ChannelRouter", "channel_router"],
        "byzpy_b200.engine.actor.ipc": ["wrap_payload", "unwrap_payload"],
        "byzpy_b200.engine.actor.backends.thread": ["ThreadActorBackend"],
        "byzpy_b200.engine.actor.backends.process": ["ProcessActorBackend"],
        "byzpy_b200.engine.actor.backends.gpu": ["GPUActorBackend", "UCXRemoteActorBackend", "UCXRemoteActorServer", "start_ucx_actor_server"],
        "byzpy_b200.engine.actor.backends.remote": ["RemoteActorBackend", "RemoteActorServer", "start_actor_server"],
        "byzpy_b200.engine.actor.transports.tcp": ["chan_put", "chan_get"],
        "byzpy_b200.engine.actor.transports.ucx": ["have_ucx"],
        "byzpy_b200.engine.node": ["NodeApplication", "NodePipeline", "HonestNodeApplication", "ByzantineNodeApplication", "CallableOp",
                                   "RemoteCallableOp", "make_single_operator_graph", "NodeContext", "InProcessContext", "ProcessContext",
                                   "RemoteContext", "DecentralizedNode", "DistributedHonestNode", "DistributedByzantineNode",
                                   "DecentralizedCluster", "MessageRouter", "RemoteNodeServer", "RemoteNodeClient", "serialize_message",
                                   "deserialize_message"],
        "byzpy_b200.engine.node.context": ["MeshRemoteContext"],
        "byzpy_b200.engine.node.actors": ["HonestNodeActor", "ByzantineNodeActor"],
        "byzpy_b200.engine.node.base": ["Node", "HonestNode", "ByzantineNode"],
        "byzpy_b200.engine.node.mixin": ["P2PHonestMixin", "P2PByzantineMixin"],
        "byzpy_b200.engine.node.remote_server": ["ServerNodeContext"],
        "byzpy_b200.engine.parameter_server.ps": ["ParameterServer"],
        "byzpy_b200.engine.parameter_server.runner": ["ParameterServerRunner"],
        "byzpy_b200.engine.parameter_server.decentralized": ["DecentralizedParameterServer"],
        "byzpy_b200.engine.peer_to_peer.topology": ["Topology", "Edge"],
        "byzpy_b200.engine.peer_to_peer.train": ["PeerToPeer"],
        "byzpy_b200.engine.peer_to_peer.runner": ["DecentralizedPeerToPeer"],
        "byzpy_b200.engine.storage.shared_store": ["SharedTensorHandle", "register_tensor", "open_tensor", "cleanup_tensor"],
        "byzpy_b200.engine.transport": ["Transport", "LocalTransport", "TcpTransport", "TcpMailbox", "send_message"],
        "byzpy_b200.engine.backend.ndarray": ["get_array_backend"],
        "byzpy_b200.utils": ["train_with_progress"],
        "byzpy_b200.cli": ["main"],
    }
    for mod, names in surface.items():
        m = importlib.import_module(mod)
        for n in names:
            assert hasattr(m, n), f"{mod}.{n} missing"


def test_row_layout_spread_handles_uneven_and_idle_ranks():
    from byzpy_b200.parallel.device_ps import RowLayout

    lay = RowLayout.spread(5, 2, 2)                       # 7 workers on 2 ranks: 4 + 3
    assert [len(lay.local_ids(r)) for r in range(2)] == [4, 3] and lay.max_local() == 4
    assert sorted(g for r in range(2) for g in lay.local_ids(r)) == list(range(7))
    idle = RowLayout.spread(6, 0, 8, n_virtual=2)        # 6 replicas on 8 ranks: two ranks host none
    assert [len(idle.local_ids(r)) for r in range(8)] == [1] * 6 + [0, 0] and idle.n_virtual == 2
    even = RowLayout.spread(6, 2, 4)
    blk = RowLayout.block(6, 2, 4)
    assert even.rank_of == blk.rank_of and even.slot_of == blk.slot_of


def test_actor_ref_attribute_calls_are_named_coroutine_functions():
    import inspect

    from byzpy_b200.engine.actor.base import ActorRef

    class Echo:
        async def start(self):
            pass

        async def close(self):
            pass

        async def call(self, method, *a, **kw):
            return (method, a, kw)

    ref = ActorRef(Echo())
    fn = ref.compute
    assert fn.__name__ == "compute" and inspect.iscoroutinefunction(fn)
    assert asyncio.run(fn(1, k=2)) == ("compute", (1,), {"k": 2})
    with pytest.raises(AttributeError):
        ref.__deepcopy__
