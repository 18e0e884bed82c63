"""Node runtime: router, contexts (in-process / process / hub / mesh), cluster, decentralized node."""
import asyncio
import socket

import pytest
import torch

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
from byzpy_b200.attacks import EmpireAttack, SignFlipAttack
from byzpy_b200.engine.graph.ops import CallableOp, make_single_operator_graph
from byzpy_b200.engine.graph.pool import ActorPoolConfig
from byzpy_b200.engine.graph.scheduler import MessageSource
from byzpy_b200.engine.node import (ByzantineNodeApplication, DecentralizedCluster, DecentralizedNode,
                                    DistributedByzantineNode, DistributedHonestNode,
                                    HonestNodeApplication, InProcessContext, MeshRemoteContext,
                                    MessageRouter, NodeApplication, ProcessContext, RemoteContext,
                                    RemoteNodeServer, deserialize_message, serialize_message)
from byzpy_b200.engine.peer_to_peer.topology import Topology


def run(coro):
    return asyncio.run(coro)


def app(name="a"):
    return NodeApplication(name=name, actor_pool=[ActorPoolConfig(backend="thread", count=1)])


def test_topology_constructors():
    t = Topology.ring(5, 1)
    assert sorted(t.out[0]) == [1, 4] and sorted(t.in_[0]) == [1, 4]
    c = Topology.complete(4)
    assert all(len(c.out[i]) == 3 for i in range(4))
    dup = Topology.ring(3, 2)
    assert len(dup.out[0]) == 4 and dup.out_neighbors(0) == [1, 2]
    with pytest.raises(ValueError):
        Topology(2, [(0, 5)])


def test_router_rules():
    topo = Topology(3, [(0, 1), (1, 2)])
    r = MessageRouter(topology=topo, node_id="a", node_id_map={0: "a", 1: "b", 2: "c"})
    assert r.get_out_neighbors() == ["b"] and r.get_in_neighbors() == []
    assert r.can_send_to("b") and not r.can_send_to("c") and not r.can_send_to("zzz")
    sent = []

    class Ctx:
        async def send_message(self, to, t, p):
            if to == "dead":
                raise RuntimeError("down")
            sent.append((to, t, p))

    async def scenario():
        await r.route_direct("b", "m", 1, Ctx())
        with pytest.raises(ValueError):
            await r.route_direct("a", "m", 1, Ctx())
        with pytest.raises(ValueError):
            await r.route_direct("c", "m", 1, Ctx())
        await r.route_broadcast("m", 2, Ctx())
        with pytest.raises(ValueError):
            await r.route_multicast(["b", "c"], "m", 3, Ctx())
        await r.route_multicast(["b"], "m", 3, Ctx())
        await r.route_reply({"from": "b"}, "m", 4, Ctx())
        with pytest.raises(ValueError):
            await r.route_reply({}, "m", 4, Ctx())
        free = MessageRouter(topology=None, node_id="x")
        assert free.can_send_to("anything") and free.get_out_neighbors() == []

    run(scenario())
    assert [s[2] for s in sent] == [1, 2, 3, 4]
    with pytest.raises(ValueError):
        MessageRouter(topology=topo, node_id=7)


def test_serialize_roundtrip():
    msg = {"from": "a", "type": "g", "payload": {"v": torch.arange(3)}}
    back = deserialize_message(serialize_message(msg))
    assert back["from"] == "a" and torch.equal(back["payload"]["v"], torch.arange(3))


def test_in_process_cluster_broadcast_pipelines_and_autonomous_tasks():
    async def scenario():
        cl = DecentralizedCluster()
        topo = Topology.ring(4, 1)
        got = {}
        for i in range(4):
            a = HonestNodeApplication(name=f"h{i}", actor_pool=[ActorPoolConfig(backend="thread", count=1)])
            a.register_pipeline("double", make_single_operator_graph(
                node_name="double", operator=CallableOp(lambda v: v * 2, input_mapping={"v": "v"}), input_keys=("v",)))
            n = await cl.add_node(node_id=f"n{i}", application=a, topology=topo, context=InProcessContext())

            async def h(frm, payload, i=i):
                got.setdefault(i, []).append((frm, payload))

            n.register_message_handler("grad", h)
        with pytest.raises(ValueError):
            await cl.add_node(node_id="n0", application=app(), context=InProcessContext())
        n0 = cl.get_node("n0")
        with pytest.raises(RuntimeError):
            await n0.send_message("n1", "grad", 1)
        await cl.start_all()
        await n0.broadcast_message("grad", {"vector": torch.ones(2)})
        await n0.multicast_message(["n1"], "grad", "mc")
        with pytest.raises(ValueError):
            await n0.send_message("n2", "grad", 1)  # not a ring neighbour
        await asyncio.sleep(0.25)
        assert sorted(got) == [1, 3] and len(got[1]) == 2
        assert (await n0.execute_pipeline("double", {"v": 21}))["double"] == 42
        with pytest.raises(KeyError):
            await n0.execute_pipeline("nope", {})
        # message-driven pipeline input
        n1 = cl.get_node("n1")
        fut = asyncio.ensure_future(n1.execute_pipeline("double", {"v": MessageSource("num")}))
        await asyncio.sleep(0.01)
        await n0.send_message("n1", "num", 4)
        assert (await fut)["double"] == 8
        ticks = []

        async def loop():
            while True:
                ticks.append(1)
                await asyncio.sleep(0.01)

        await n0.start_autonomous_task(loop(), "ticker")
        with pytest.raises(ValueError):
            await n0.start_autonomous_task(asyncio.sleep(0), "ticker")
        await asyncio.sleep(0.05)
        await cl.remove_node("n3")
        assert cl.get_node("n3") is None
        await cl.shutdown_all()
        assert ticks and not InProcessContext._registry

    run(scenario())
    with pytest.raises(ValueError):
        DecentralizedNode(node_id="", application=app(), context=InProcessContext())


def test_process_context_two_nodes():
    async def scenario():
        cl = DecentralizedCluster()
        topo = Topology.complete(2)
        got = []
        for i in range(2):
            n = await cl.add_node(node_id=f"p{i}", application=app(f"a{i}"), topology=topo)  # ProcessContext
            assert isinstance(n.context, ProcessContext)

            async def h(frm, payload, i=i):
                got.append((i, frm, payload))

            n.register_message_handler("hello", h)
        await cl.start_all()
        await cl.get_node("p0").send_message("p1", "hello", {"t": torch.arange(3)})
        await cl.get_node("p1").send_message("p0", "hello", "pong")
        for _ in range(50):
            if len(got) >= 2:
                break
            await asyncio.sleep(0.05)
        await cl.shutdown_all()
        assert sorted((g[0], g[1]) for g in got) == [(0, "p1"), (1, "p0")]

    run(scenario())


def test_hub_server_and_remote_context():
    async def scenario():
        srv = RemoteNodeServer("127.0.0.1", 0)
        await srv.start()
        got = []
        hosted = DecentralizedNode(node_id="hosted", application=app("s"), context=InProcessContext())

        async def h(frm, payload):
            got.append(("hosted", frm, payload))

        hosted.register_message_handler("m", h)
        await srv.register_node(hosted)
        clients = []
        for name in ("c1", "c2"):
            c = DecentralizedNode(node_id=name, application=app(name), context=RemoteContext("127.0.0.1", srv.port))

            async def hc(frm, payload, name=name):
                got.append((name, frm, payload))

            c.register_message_handler("m", hc)
            await c.start()
            clients.append(c)
        await clients[0].send_message("hosted", "m", 1)
        await clients[0].send_message("c2", "m", 2)     # relayed client -> client
        await hosted.send_message("c1", "m", 3)
        await asyncio.sleep(0.4)
        assert sorted(got) == [("c1", "hosted", 3), ("c2", "c1", 2), ("hosted", "c1", 1)]
        for c in clients:
            await c.shutdown()
        await srv.shutdown()
        with pytest.raises(ConnectionError):
            bad = DecentralizedNode(node_id="x", application=app(), context=RemoteContext("127.0.0.1", 1))
            await bad.start()

    run(scenario())


def _ports(k):
    out = []
    for _ in range(k):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        out.append(s.getsockname()[1])
        s.close()
    return out


def test_mesh_context_full_mesh_and_reconnect():
    async def scenario():
        ports = _ports(3)
        addrs = {f"m{i}": ("127.0.0.1", ports[i]) for i in range(3)}
        idmap = {i: f"m{i}" for i in range(3)}
        topo = Topology.complete(3)
        got = []

        def make(i):
            ctx = MeshRemoteContext("127.0.0.1", ports[i], {k: v for k, v in addrs.items() if k != f"m{i}"},
                                    reconnect_interval=0.2)
            n = DecentralizedNode(node_id=f"m{i}", application=app(f"a{i}"), context=ctx, topology=topo,
                                  node_id_map=idmap)

            async def h(frm, payload, i=i):
                got.append((i, frm, payload))

            n.register_message_handler("x", h)
            return n

        nodes = [make(i) for i in range(3)]
        await nodes[0].start()
        await nodes[1].start()
        await asyncio.sleep(0.1)
        await nodes[2].start()          # late joiner: earlier nodes re-dial it
        await asyncio.sleep(0.6)
        await nodes[0].broadcast_message("x", "hi")
        await nodes[2].send_message("m0", "x", "direct")
        await asyncio.sleep(0.4)
        assert sorted(got) == [(0, "m2", "direct"), (1, "m0", "hi"), (2, "m0", "hi")]
        assert set(nodes[0].context.get_connected_peers()) == {"m1", "m2"}
        # kill m1 and bring a replacement up on the same port: the monitor reconnects
        await nodes[1].shutdown()
        await asyncio.sleep(0.3)
        nodes[1] = make(1)
        await nodes[1].start()
        await asyncio.sleep(0.8)
        got.clear()
        await nodes[0].send_message("m1", "x", "again")
        await asyncio.sleep(0.3)
        assert got == [(1, "m0", "again")]
        for n in nodes:
            await n.shutdown()

    run(scenario())


class _Hon(DistributedHonestNode):
    def __init__(self):
        super().__init__(actor_pool=[ActorPoolConfig(backend="thread", count=2)], aggregator=CoordinateWiseMedian())

    def next_batch(self):
        return torch.ones(2), torch.zeros(2)

    def apply_server_gradient(self, g):
        self.last = g

    def local_honest_gradient(self, *, x, y):
        return x * 3


class _Byz(DistributedByzantineNode):
    def __init__(self, attack):
        super().__init__(actor_pool=[ActorPoolConfig(backend="thread", count=1)], attack=attack)

    def next_batch(self):
        return torch.empty(0), torch.empty(0)

    def apply_server_gradient(self, g):
        pass


class _CustomByz(DistributedByzantineNode):
    def __init__(self):
        super().__init__(actor_pool=[ActorPoolConfig(backend="thread", count=2)])

    def next_batch(self):
        return torch.empty(0), torch.empty(0)

    def apply_server_gradient(self, g):
        pass

    def byzantine_gradient(self, x, y, honest_grads=None):
        return -10 * torch.stack(list(honest_grads)).mean(0)


def test_distributed_nodes_pipelines():
    h = _Hon()
    assert torch.equal(h.honest_gradient_for_next_batch(), torch.full((2,), 3.0))
    g = [torch.tensor([1.0]), torch.tensor([5.0]), torch.tensor([9.0])]
    assert h.aggregate_sync(g).item() == 5.0
    assert run(h.aggregate(g)).item() == 5.0
    b = _Byz(EmpireAttack(scale=-1.0))
    assert torch.allclose(b.byzantine_gradient_for_next_batch(g), torch.tensor([-5.0]))
    with pytest.raises(ValueError):
        _Byz(SignFlipAttack()).byzantine_gradient_for_next_batch(g)  # needs base_grad: subclass must supply it
    c = _CustomByz()
    assert torch.allclose(c.byzantine_gradient_for_next_batch(g), torch.tensor([-50.0]))
    with pytest.raises(RuntimeError):
        c.prepare_attack_inputs(honest_grads=g)
    with pytest.raises(ValueError):
        DistributedByzantineNode.__init__(_Byz.__new__(_Byz), actor_pool=[ActorPoolConfig("thread")])
    for n in (h, b, c):
        run(n.shutdown_distributed())


def test_application_reserved_pipelines_and_sync_guard():
    a = HonestNodeApplication(name="x", actor_pool=[ActorPoolConfig("thread")])
    with pytest.raises(KeyError):
        a.aggregate_sync(gradients=[])
    g = make_single_operator_graph(node_name="aggregate", operator=CoordinateWiseMedian(), input_keys=("gradients",))
    a.register_pipeline("aggregate", g)
    with pytest.raises(ValueError):
        a.register_pipeline("aggregate", g)
    assert list(a.list_pipelines()) == ["aggregate"] and a.has_pipeline("aggregate")

    async def inside():
        with pytest.raises(RuntimeError):
            a.run_pipeline_sync("aggregate", {"gradients": [torch.ones(1)]})
        with pytest.raises(KeyError):
            await a.run_pipeline("zzz", {})

    run(inside())
    b = ByzantineNodeApplication(name="b", actor_pool=[ActorPoolConfig("thread")])
    with pytest.raises(KeyError):
        b.run_attack_sync(inputs={})
    run(a.shutdown())
    run(b.shutdown())
