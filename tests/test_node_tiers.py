"""Fine-grained tiers of the node runtime, one behaviour per test: topology, message router,
wire framing, hub client/server, node applications, decentralized node lifecycle / handlers /
pipelines / autonomous tasks, in-process context and cluster bookkeeping (mirrors reference
tests/engine/node/test_{router,application,decentralized,cluster,context,remote_*}.py and
tests/engine/peer_to_peer/test_topology.py)."""
import asyncio

import pytest
import torch

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
from byzpy_b200.attacks import EmpireAttack
from byzpy_b200.engine.graph.graph import ComputationGraph, GraphNode, graph_input
from byzpy_b200.engine.graph.operator import MessageTriggerOp, Operator
from byzpy_b200.engine.graph.ops import CallableOp, make_single_operator_graph
from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig
from byzpy_b200.engine.graph.scheduler import MessageSource
from byzpy_b200.engine.node import (ByzantineNodeApplication, DecentralizedCluster, DecentralizedNode,
                                    HonestNodeApplication, InProcessContext, MessageRouter, NodeApplication,
                                    NodePipeline, RemoteNodeClient, RemoteNodeServer, deserialize_message,
                                    serialize_message)
from byzpy_b200.engine.node.remote_client import read_frame, write_frame
from byzpy_b200.engine.peer_to_peer.topology import Edge, Topology


def run(coro):
    return asyncio.run(coro)


def _pool():
    return [ActorPoolConfig(backend="thread", count=1)]


def _app(name="n", cls=NodeApplication, **kw):
    return cls(name=name, actor_pool=_pool(), **kw)


def _fn_graph(fn, keys=("x",), node="out"):
    op = CallableOp(fn, input_mapping={k: k for k in keys})
    return make_single_operator_graph(node_name=node, operator=op, input_keys=keys)


class _Ctx:
    """Context double that records sends; ids listed in ``dead`` raise."""

    def __init__(self, dead=()):
        self.sent, self.dead = [], set(dead)

    async def send_message(self, to, mtype, payload):
        if to in self.dead:
            raise ConnectionError(f"{to} is down")
        self.sent.append((to, mtype, payload))


# ------------------------------------------------------------------------------------------ topology
@pytest.mark.parametrize("n,k,degree", [(4, 1, 2), (6, 2, 4), (5, 1, 2), (2, 1, 1), (3, 1, 2)])
def test_ring_degrees(n, k, degree):
    t = Topology.ring(n, k)
    assert all(len(t.out_neighbors(i)) == degree and len(t.in_neighbors(i)) == degree for i in range(n))
    assert all(i not in t.out_neighbors(i) for i in range(n))


def test_ring_is_symmetric_and_keeps_duplicates_raw():
    t = Topology.ring(4, 2)
    assert all((v, u) in set(t.edges()) for u, v in t.edges())
    assert len(t.out[0]) == 4 and t.out_neighbors(0) == [1, 3, 2] and t.out_neighbors(0, unique=False) == [1, 3, 2, 2]


@pytest.mark.parametrize("n", [1, 2, 5])
def test_complete_topology(n):
    t = Topology.complete(n)
    assert t.n == n and len(t.edges()) == n * (n - 1)
    assert all(sorted(t.out[i]) == [j for j in range(n) if j != i] == sorted(t.in_[i]) for i in range(n))


def test_directed_edges_and_edge_type():
    t = Topology(3, [(0, 1), (0, 2), (2, 1)])
    assert t.out == {0: [1, 2], 1: [], 2: [1]} and t.in_ == {0: [], 1: [0, 2], 2: [0]}
    e = t.edges()[0]
    assert isinstance(e, Edge) and (e.u, e.v) == (0, 1) and e == (0, 1)


@pytest.mark.parametrize("bad", [(-1, 0), (0, 3), (3, 3)])
def test_topology_rejects_out_of_range_edges(bad):
    with pytest.raises(ValueError, match="outside"):
        Topology(3, [bad])


# -------------------------------------------------------------------------------------------- router
def _line_router(me="b"):
    return MessageRouter(topology=Topology(3, [(0, 1), (1, 2), (1, 0)]), node_id=me,
                         node_id_map={0: "a", 1: "b", 2: "c"})


def test_router_id_translation():
    r = _line_router()
    assert r._to_internal_id("c") == 2 and r._to_internal_id(1) == 1 and r._to_internal_id("ghost") == -1
    assert r._to_external_id(0) == "a" and r._to_external_id(9) == 9


def test_router_neighbours_with_string_and_integer_ids():
    r = _line_router("b")
    assert r.get_out_neighbors() == ["c", "a"] and r.get_in_neighbors() == ["a"]
    ints = MessageRouter(topology=Topology.ring(4, 1), node_id=2)
    assert ints.get_out_neighbors() == [3, 1] and ints.get_out_neighbors_internal() == [3, 1]
    assert ints.can_send_to(3) and not ints.can_send_to(0) and not ints.can_send_to(2)


def test_router_validates_integer_node_id_range():
    with pytest.raises(ValueError, match="not in topology"):
        MessageRouter(topology=Topology.ring(3, 1), node_id=3)
    with pytest.raises(ValueError):
        MessageRouter(topology=Topology.ring(3, 1), node_id=-1)
    MessageRouter(topology=Topology.ring(3, 1), node_id="anything")       # strings are checked lazily


def test_router_direct_rules():
    r, ctx = _line_router("a"), _Ctx()

    async def go():
        await r.route_direct("b", "g", 1, ctx)
        await r.route_message("b", "g", 2, ctx)
        with pytest.raises(ValueError, match="self"):
            await r.route_direct("a", "g", 0, ctx)
        with pytest.raises(ValueError, match="not a neighbor"):
            await r.route_direct("c", "g", 0, ctx)

    run(go())
    assert ctx.sent == [("b", "g", 1), ("b", "g", 2)]


def test_router_broadcast_dedups_and_survives_dead_neighbours():
    r = MessageRouter(topology=Topology.ring(3, 2), node_id=0)      # raw out list is [1, 2, 2, 1]
    ctx = _Ctx(dead={1})
    run(r.route_broadcast("m", "p", ctx))
    assert ctx.sent == [(2, "m", "p")]
    run(r.route_broadcast("m", "p", None))                          # no context: silently nothing


def test_router_multicast_is_all_or_nothing():
    r, ctx = _line_router("b"), _Ctx()

    async def go():
        await r.route_multicast([], "m", 0, ctx)
        with pytest.raises(ValueError):
            await r.route_multicast(["a", "zzz"], "m", 0, ctx)
        assert ctx.sent == []                                       # validated before the first send
        await r.route_multicast(["a", "c"], "m", 1, ctx)
        await r.route_multicast(["a"], "m", 2, None)

    run(go())
    assert ctx.sent == [("a", "m", 1), ("c", "m", 1)]


def test_router_reply_goes_to_sender_and_obeys_topology():
    ctx = _Ctx()

    async def go():
        await _line_router("b").route_reply({"from": "a", "type": "q"}, "ans", 42, ctx)
        with pytest.raises(ValueError, match="no 'from'"):
            await _line_router("b").route_reply({"type": "q"}, "ans", 42, ctx)
        with pytest.raises(ValueError, match="not a neighbor"):
            await _line_router("c").route_reply({"from": "a"}, "ans", 42, ctx)   # c has no out-edges

    run(go())
    assert ctx.sent == [("a", "ans", 42)]


def test_router_without_topology_allows_everything_but_self():
    r, ctx = MessageRouter(topology=None, node_id="me"), _Ctx()

    async def go():
        await r.route_direct("anyone", "m", 1, ctx)
        with pytest.raises(ValueError):
            await r.route_direct("me", "m", 1, ctx)
        await r.route_broadcast("m", 2, ctx)                        # nobody to broadcast to
        await r.route_multicast(["x", "y"], "m", 3, ctx)

    run(go())
    assert [s[0] for s in ctx.sent] == ["anyone", "x", "y"]
    assert r.get_in_neighbors() == [] and r.get_out_neighbors_internal() == []


# ---------------------------------------------------------------------------------------- wire format
def test_serialize_tags_and_legacy_frames():
    import cloudpickle

    msg = {"from": "a", "type": "t", "payload": [1, {"k": (2, 3)}]}
    data = serialize_message(msg)
    assert data[:1] == b"P" and deserialize_message(data) == msg
    assert deserialize_message(cloudpickle.dumps(msg)) == msg        # untagged legacy body


def test_serialize_tensors_and_nested_containers():
    payload = {"g": torch.arange(6.0).reshape(2, 3), "list": [torch.ones(2, dtype=torch.int64)], "meta": ("x", 1.5)}
    back = deserialize_message(serialize_message({"payload": payload}))["payload"]
    assert torch.equal(back["g"], payload["g"]) and back["list"][0].dtype == torch.int64 and back["meta"] == ("x", 1.5)


def test_serialize_gpu_direct_falls_back_on_cpu_only_hosts():
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    data = serialize_message({"payload": torch.ones(2)}, gpu_direct=True)
    assert data[:1] == b"P" and torch.equal(deserialize_message(data)["payload"], torch.ones(2))


def test_frame_roundtrip_over_a_socket_pair():
    async def go():
        got = []

        async def handler(reader, writer):
            got.append(await read_frame(reader))
            got.append(await read_frame(reader))
            await write_frame(writer, {"ack": len(got)})
            writer.close()

        server = await asyncio.start_server(handler, "127.0.0.1", 0)
        port = server.sockets[0].getsockname()[1]
        reader, writer = await asyncio.open_connection("127.0.0.1", port)
        await write_frame(writer, {"i": 1, "t": torch.zeros(1000)})
        await write_frame(writer, {"i": 2})
        assert await read_frame(reader) == {"ack": 2}
        with pytest.raises(asyncio.IncompleteReadError):
            await read_frame(reader)                                 # peer closed
        writer.close()
        server.close()
        await server.wait_closed()
        return got

    got = run(go())
    assert got[0]["i"] == 1 and got[0]["t"].numel() == 1000 and got[1] == {"i": 2}


# -------------------------------------------------------------------------------- hub client / server
def test_client_connect_errors_and_state():
    async def go():
        c = RemoteNodeClient("127.0.0.1", 1)
        assert not c.is_connected()
        with pytest.raises(ConnectionError):
            await c.connect(timeout=1.0)
        with pytest.raises(RuntimeError, match="not connected"):
            await c.send_message("x", "t", 1)
        with pytest.raises(RuntimeError, match="Failed to register"):
            await c.register_node("me")
        assert await c.receive_message(timeout=0.01) is None
        await c.disconnect()                                         # harmless when never connected

    run(go())


def test_hub_relays_between_clients_and_drops_unknown_targets():
    async def go():
        server = RemoteNodeServer("127.0.0.1", 0)
        await server.start()
        a, b = RemoteNodeClient("127.0.0.1", server.port), RemoteNodeClient("127.0.0.1", server.port)
        for c, nid in ((a, "a"), (b, "b")):
            await c.connect()
            await c.connect()                                        # idempotent
            await c.register_node(nid)
        await asyncio.sleep(0.05)
        await a.send_message("nobody", "lost", 0, from_node_id="a")  # server swallows routing errors
        await a.send_message("b", "grad", {"v": torch.ones(3)}, from_node_id="a")
        msg = await b.receive_message(timeout=2.0)
        assert msg["from"] == "a" and msg["type"] == "grad" and msg["to"] == "b"
        assert torch.equal(msg["payload"]["v"], torch.ones(3))
        await b.send_message("a", "ack", 1)                          # no from -> "unknown"
        assert (await a.receive_message(timeout=2.0))["from"] == "unknown"
        await b.disconnect()
        await asyncio.sleep(0.05)
        assert "b" not in server._clients and not b.is_connected()
        with pytest.raises(ValueError, match="not found"):
            await server.send_message_to_client("b", {"type": "x"})
        await a.disconnect()
        await server.shutdown()

    run(go())


def test_hub_delivers_to_server_side_nodes_and_rejects_duplicates():
    async def go():
        server = RemoteNodeServer("127.0.0.1", 0)
        await server.start()
        seen = []
        node = DecentralizedNode(node_id="hub-node", application=_app(), context=InProcessContext())

        async def on_ping(frm, payload):
            seen.append((frm, payload))

        node.register_message_handler("ping", on_ping)
        await server.register_node(node)
        with pytest.raises(ValueError, match="already registered"):
            await server.register_node(node)
        c = RemoteNodeClient("127.0.0.1", server.port)
        await c.connect()
        await c.register_node("remote")
        await c.send_message("hub-node", "ping", 7, from_node_id="remote")
        for _ in range(100):
            if seen:
                break
            await asyncio.sleep(0.01)
        await node.send_message("remote", "pong", 8)                 # server-side node -> TCP client
        msg = await c.receive_message(timeout=2.0)
        await c.disconnect()
        await server.shutdown()
        return seen, msg

    seen, msg = run(go())
    assert seen == [("remote", 7)] and msg["type"] == "pong" and msg["payload"] == 8 and msg["from"] == "hub-node"


def test_client_notices_server_shutdown():
    async def go():
        server = RemoteNodeServer("127.0.0.1", 0)
        await server.start()
        c = RemoteNodeClient("127.0.0.1", server.port)
        await c.connect()
        await c.register_node("x")
        await asyncio.sleep(0.02)
        await server.shutdown()
        for _ in range(100):
            if not c._connected:
                break
            await asyncio.sleep(0.01)
        assert not c.is_connected()
        with pytest.raises(RuntimeError):
            await c.send_message("y", "t", 1)
        await c.disconnect()

    run(go())


# --------------------------------------------------------------------------------------- applications
def test_application_pipeline_registry():
    app = _app("alpha")
    g = _fn_graph(lambda x: x + 1)
    app.register_pipeline("inc", g, metadata={"k": 1})
    assert app.has_pipeline("inc") and not app.has_pipeline("dec") and list(app.list_pipelines()) == ["inc"]
    assert isinstance(app._pipelines["inc"], NodePipeline) and app._pipelines["inc"].graph is g
    with pytest.raises(ValueError, match="already registered"):
        app.register_pipeline("inc", g)
    with pytest.raises(KeyError, match="Unknown pipeline 'dec' for node 'alpha'"):
        run(app.run_pipeline("dec", {}))
    assert isinstance(app.pool, ActorPool)


def test_application_accepts_an_existing_pool():
    pool = ActorPool(_pool())
    assert NodeApplication(name="n", actor_pool=pool).pool is pool


def test_application_metadata_layers_override_in_order():
    seen = {}

    class Peek(Operator):
        name = "peek"

        def compute(self, inputs, *, context):
            seen.clear()
            seen.update({k: v for k, v in context.metadata.items() if not k.startswith(("pool", "worker"))})
            return 0

    app = _app("n1", metadata={"a": "base", "b": "base", "c": "base"})
    app.register_pipeline("p", ComputationGraph([GraphNode("o", Peek(), {})]), metadata={"b": "pipe", "c": "pipe"})
    app.run_pipeline_sync("p", {}, metadata={"c": "call"})
    assert seen == {"node": "n1", "pipeline": "p", "a": "base", "b": "pipe", "c": "call"}
    run(app.shutdown())


def test_application_sync_helpers_refuse_running_loops():
    app = _app()
    app.register_pipeline("inc", _fn_graph(lambda x: x + 1))
    assert app.run_pipeline_sync("inc", {"x": 1}) == {"out": 2}

    async def inside():
        with pytest.raises(RuntimeError, match="async context"):
            app.run_pipeline_sync("inc", {"x": 1})
        assert await app.run_pipeline("inc", {"x": 2}) == {"out": 3}

    run(inside())
    run(app.shutdown())


def test_honest_application_reserved_pipelines():
    app = _app("h", HonestNodeApplication)
    vs = [torch.randn(9) for _ in range(5)]
    with pytest.raises(KeyError, match="No aggregation pipeline"):
        app.aggregate_sync(gradients=vs)
    with pytest.raises(KeyError, match="No honest gradient pipeline"):
        run(app.honest_gradient({"x": 1}))
    app.register_pipeline(HonestNodeApplication.AGGREGATION_PIPELINE,
                          make_single_operator_graph(node_name="agg", operator=CoordinateWiseMedian(),
                                                     input_keys=("gradients",)))
    app.register_pipeline(HonestNodeApplication.GRADIENT_PIPELINE, _fn_graph(lambda x, y: x * y, ("x", "y")))
    assert torch.equal(app.aggregate_sync(gradients=vs), CoordinateWiseMedian().aggregate(vs))
    assert torch.equal(run(app.aggregate(gradients=vs)), CoordinateWiseMedian().aggregate(vs))
    assert app.honest_gradient_sync({"x": 3, "y": 4}) == 12 and run(app.honest_gradient({"x": 2, "y": 2})) == 4
    run(app.shutdown())


def test_byzantine_application_attack_pipeline():
    app = _app("b", ByzantineNodeApplication)
    with pytest.raises(KeyError, match="No attack pipeline"):
        app.run_attack_sync(inputs={})
    app.register_pipeline(ByzantineNodeApplication.ATTACK_PIPELINE,
                          make_single_operator_graph(node_name="atk", operator=EmpireAttack(scale=-2.0),
                                                     input_keys=("honest_grads",)))
    vs = [torch.ones(3), 3 * torch.ones(3)]
    assert torch.allclose(app.run_attack_sync(inputs={"honest_grads": vs}), torch.full((3,), -4.0))
    assert torch.allclose(run(app.run_attack(inputs={"honest_grads": vs})), torch.full((3,), -4.0))
    run(app.shutdown())


# ------------------------------------------------------------------------------------ decentralized node
def _node(nid="n0", topology=None, id_map=None, app=None):
    return DecentralizedNode(node_id=nid, application=app or _app(str(nid)), context=InProcessContext(),
                             topology=topology, node_id_map=id_map)


@pytest.mark.parametrize("bad", [None, ""])
def test_node_rejects_empty_ids(bad):
    with pytest.raises(ValueError, match="cannot be empty"):
        _node(bad)


def test_node_id_zero_is_valid_and_lands_in_scheduler_metadata():
    n = _node(0)
    assert n.node_id == 0 and n.scheduler.metadata["node_id"] == 0 and n.scheduler.pool is n.application.pool


def test_node_requires_start_for_io_and_pipelines():
    n = _node()

    async def go():
        for coro in (n.send_message("x", "t", 1), n.broadcast_message("t", 1), n.multicast_message(["x"], "t", 1),
                     n.execute_pipeline("p", {})):
            with pytest.raises(RuntimeError, match="not started"):
                await coro
        idle = asyncio.sleep(0)
        with pytest.raises(RuntimeError, match="must be started"):
            await n.start_autonomous_task(idle)
        with pytest.raises(RuntimeError, match="already awaited|closed"):
            await idle                  # the refused coroutine was closed, not left to warn at GC

    run(go())


def test_node_start_and_shutdown_are_idempotent():
    async def go():
        n = _node("solo")
        await n.start()
        task = n._message_task
        await n.start()
        assert n._message_task is task and InProcessContext._registry["solo"] is n.context
        await n.shutdown()
        await n.shutdown()
        assert "solo" not in InProcessContext._registry and task.done()

    run(go())


def test_node_handlers_receive_sender_and_payload_and_can_be_replaced():
    async def go():
        topo = Topology.complete(2)
        ids = {0: "a", 1: "b"}
        a, b = _node("a", topo, ids), _node("b", topo, ids)
        got = []

        async def first(frm, payload):
            got.append(("first", frm, payload))

        async def second(frm, payload):
            got.append(("second", frm, payload))

        b.register_message_handler("g", first)
        await a.start()
        await b.start()
        await a.send_message("b", "g", 1)
        await asyncio.sleep(0.05)
        b.register_message_handler("g", second)
        await a.send_message("b", "g", 2)
        await a.send_message("b", "unhandled", 3)                    # no handler: only queued for pipelines
        await asyncio.sleep(0.05)
        assert await b.scheduler.wait_for_message("unhandled", timeout=0.5) == 3
        await a.shutdown()
        await b.shutdown()
        return got

    assert run(go()) == [("first", "a", 1), ("second", "a", 2)]


def test_node_send_respects_topology_and_neighbour_queries():
    async def go():
        topo = Topology(3, [(0, 1), (1, 2)])
        ids = {0: "a", 1: "b", 2: "c"}
        nodes = [_node(i, topo, ids) for i in "abc"]
        for n in nodes:
            await n.start()
        a, b, c = nodes
        assert a.get_neighbors() == ["b"] and b.get_in_neighbors() == ["a"] and c.get_neighbors() == []
        with pytest.raises(ValueError, match="not a neighbor"):
            await a.send_message("c", "t", 1)
        with pytest.raises(ValueError):
            await c.multicast_message(["a"], "t", 1)
        await a.multicast_message(["b"], "t", 5)
        await c.broadcast_message("t", 6)                            # no out-neighbours: no-op
        await asyncio.sleep(0.05)
        assert await b.scheduler.wait_for_message("t", timeout=0.5) == 5
        for n in nodes:
            await n.shutdown()

    run(go())


def test_node_pipeline_waits_for_messages():
    async def go():
        topo = Topology.complete(2)
        ids = {0: "a", 1: "b"}
        a, b = _node("a", topo, ids), _node("b", topo, ids)
        g = ComputationGraph([GraphNode("wait", MessageTriggerOp("go"), {}),
                              GraphNode("sum", CallableOp(lambda trig, extra, base: trig + extra + base,
                                                          input_mapping={"trig": "trig", "extra": "extra", "base": "base"}),
                                        {"trig": "wait", "extra": MessageSource("bonus", field="v"), "base": graph_input("base")})])
        b.application.register_pipeline("gather", g)
        await a.start()
        await b.start()
        pending = asyncio.ensure_future(b.execute_pipeline("gather", {"base": 100}))
        await asyncio.sleep(0.02)
        assert not pending.done()
        await a.send_message("b", "go", 1)
        await a.send_message("b", "bonus", {"v": 10})
        out = await asyncio.wait_for(pending, 2.0)
        with pytest.raises(KeyError, match="Unknown pipeline"):
            await b.execute_pipeline("nope", {})
        await a.shutdown()
        await b.shutdown()
        return out

    assert run(go()) == {"sum": 111}


def test_node_autonomous_tasks_are_named_and_cancelled_on_shutdown():
    async def go():
        n = _node("auto")
        await n.start()
        ticks = []

        async def ticker():
            while True:
                ticks.append(1)
                await asyncio.sleep(0.005)

        task = await n.start_autonomous_task(ticker(), name="tick")
        dup = ticker()
        with pytest.raises(ValueError, match="already exists"):
            await n.start_autonomous_task(dup, name="tick")
        dup.close()
        other = await n.start_autonomous_task(asyncio.sleep(0), name="other")
        await asyncio.sleep(0.03)
        await n.shutdown()
        assert task.cancelled() and other.done() and len(ticks) >= 2 and n._autonomous_tasks == {}

    run(go())


def test_handler_exception_does_not_lose_scheduler_delivery():
    async def go():
        n = _node("h")

        async def bad(frm, payload):
            raise RuntimeError("handler bug")

        n.register_message_handler("t", bad)
        with pytest.raises(RuntimeError, match="handler bug"):
            await n.handle_incoming_message("x", "t", 9)
        assert await n.scheduler.wait_for_message("t", timeout=0.1) == 9

    run(go())


# --------------------------------------------------------------------------------- in-process context
def test_in_process_context_rules():
    async def go():
        a, b = _node("ctx-a"), _node("ctx-b")
        with pytest.raises(RuntimeError, match="not started"):
            await a.context.send_message("ctx-b", "t", 1)
        await a.start()
        with pytest.raises(ValueError, match="not found or not running"):
            await a.context.send_message("ctx-b", "t", 1)
        await b.start()
        b._message_task.cancel()                                     # keep the node from draining its inbox
        t = torch.ones(4)
        await a.context.send_message("ctx-b", "t", t)
        msg = await asyncio.wait_for(b.context._inbox.get(), 1.0)
        assert msg == {"from": "ctx-a", "type": "t", "payload": t} and msg["payload"] is t      # by reference
        await a.shutdown()
        await b.shutdown()
        assert b.context._inbox.empty() and "ctx-b" not in InProcessContext._registry

    run(go())


# ------------------------------------------------------------------------------------------- cluster
def test_cluster_bookkeeping_and_id_maps():
    async def go():
        cl = DecentralizedCluster()
        topo = Topology.ring(3, 1)
        nodes = [await cl.add_node(node_id=f"n{i}", application=_app(f"n{i}"), topology=topo,
                                   context=InProcessContext()) for i in range(3)]
        with pytest.raises(ValueError, match="already exists"):
            await cl.add_node(node_id="n0", application=_app(), context=InProcessContext())
        assert cl.get_node("n1") is nodes[1] and cl.get_node("zz") is None
        assert all(n.message_router._node_id_map == {0: "n0", 1: "n1", 2: "n2"} for n in nodes)   # late joiners propagate
        assert nodes[0].get_neighbors() == ["n1", "n2"]
        await cl.start_all()
        got = []

        async def on(frm, payload):
            got.append((frm, payload))

        nodes[1].register_message_handler("hi", on)
        await nodes[0].broadcast_message("hi", "x")
        await asyncio.sleep(0.05)
        await cl.remove_node("n0")
        await cl.remove_node("n0")                                   # unknown id: no-op
        assert cl._node_id_map == {0: "n1", 1: "n2"} and nodes[1].message_router._reverse_id_map == {"n1": 0, "n2": 1}
        await cl.shutdown_all()
        assert cl.nodes == {} and cl._node_id_map == {}
        return got

    assert run(go()) == [("n0", "x")]


def test_cluster_defaults_to_process_context():
    from byzpy_b200.engine.node import ProcessContext

    async def go():
        cl = DecentralizedCluster()
        node = await cl.add_node(node_id="p", application=_app("p"))
        assert isinstance(node.context, ProcessContext)              # never started: nothing spawned

    run(go())


def test_a_raising_message_handler_does_not_make_the_node_deaf():
    import warnings

    from byzpy_b200.engine.node.application import HonestNodeApplication
    from byzpy_b200.engine.node.context import InProcessContext
    from byzpy_b200.engine.node.decentralized import DecentralizedNode
    from byzpy_b200.engine.graph.pool import ActorPoolConfig

    async def scenario():
        nodes = [DecentralizedNode(node_id=i, application=HonestNodeApplication(name=f"n{i}", actor_pool=[ActorPoolConfig(backend="thread", count=1)]),
                                   context=InProcessContext()) for i in range(2)]
        seen = []

        async def flaky(frm, payload):
            if payload == "boom":
                raise ValueError("handler bug")
            seen.append(payload)

        nodes[1].register_message_handler("note", flaky)
        for n in nodes:
            await n.start()
        try:
            with warnings.catch_warnings(record=True) as caught:
                warnings.simplefilter("always")
                await nodes[0].send_message(1, "note", "one")
                await nodes[0].send_message(1, "note", "boom")
                await nodes[0].send_message(1, "note", "two")
                await nodes[0].send_message(1, "note", "boom")
                await nodes[0].send_message(1, "note", "three")
                for _ in range(100):
                    if len(seen) == 3:
                        break
                    await asyncio.sleep(0.02)
            assert seen == ["one", "two", "three"]
            assert [t for t, _ in nodes[1].handler_errors] == ["note", "note"]
            assert sum("handler of 'note' messages raised" in str(w.message) for w in caught) == 1
        finally:
            for n in nodes:
                await n.shutdown()

    asyncio.run(scenario())
