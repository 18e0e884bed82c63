"""Message transports behind a decentralized node (reference engine/node/context.py:11-1064).

``NodeContext`` is the four-method contract (``start`` / ``send_message`` / ``receive_messages``
/ ``shutdown``); messages are dicts ``{"from", "type", "payload"}``.

* ``InProcessContext``  -- class-level registry + asyncio queues; tensors (CPU or CUDA) are handed
  over by reference, i.e. zero-copy between device-resident nodes in one process.
* ``ProcessContext``    -- the node runs as a mirror ``DecentralizedNode`` inside a spawned OS
  process (its ``init_callback`` executes there, so models can be built in the child); parent and
  child talk over a duplex pipe serviced by reader threads that wake the event loop directly
  (no polling sleeps; the child signals readiness instead of a fixed start-up delay).  The parent
  relays child->child traffic through the registry.
* ``RemoteContext``     -- client of a :class:`RemoteNodeServer` hub.
* ``MeshRemoteContext`` -- every node is a TCP server plus clients to all peers; dead peers are
  re-dialled every ``reconnect_interval`` seconds and sends fall back outbound -> inbound.
  ``gpu_direct=True`` ships CUDA tensors as CUDA IPC handles between processes on one NVSwitch box.
"""
from __future__ import annotations

import asyncio
import threading
from abc import ABC, abstractmethod
from typing import TYPE_CHECKING, Any, AsyncIterator, Dict, Optional, Tuple

import cloudpickle

from ..actor.backends.process import process_context

if TYPE_CHECKING:  # pragma: no cover
    from .decentralized import DecentralizedNode


class PeerUnreachableError(ConnectionError, RuntimeError):
    """A message could not be handed to a peer: no inbound connection from it and dialling it failed.
    It is a ``ConnectionError`` (transport layer) and, like the reference's error, a ``RuntimeError``."""


class NodeContext(ABC):
    """Where a decentralized node lives and how its messages travel: the transport contract.

    Four methods: ``await start(node)`` attaches the context to its :class:`~byzpy_b200.engine.node.decentralized.
    DecentralizedNode`; ``await send_message(to_node_id, message_type, payload)``; ``receive_messages()`` is an async
    iterator of ``{"from", "type", "payload"}`` dicts that ends when the context shuts down; ``await shutdown()``.
    Implementations: :class:`InProcessContext`, :class:`ProcessContext`, :class:`RemoteContext`,
    :class:`MeshRemoteContext`.
    """

    @abstractmethod
    async def start(self, node: "DecentralizedNode") -> None:
        """Attach to ``node`` and become able to send and receive."""

    @abstractmethod
    async def send_message(self, to_node_id: Any, message_type: str, payload: Any) -> None:
        """Deliver ``{"from": this node, "type": message_type, "payload": payload}`` to ``to_node_id``."""

    @abstractmethod
    def receive_messages(self) -> AsyncIterator[Any]:
        """Async iterator over incoming message dicts; ends when the context is shut down."""

    @abstractmethod
    async def shutdown(self) -> None:
        """Stop receiving and release the transport."""


async def _drain(queue: "asyncio.Queue", is_running, idle: float = 0.1) -> AsyncIterator[Any]:
    """Yield queue items while ``is_running()``; wakes up every ``idle`` seconds to re-check."""
    while is_running():
        try:
            yield await asyncio.wait_for(queue.get(), timeout=idle)
        except asyncio.TimeoutError:
            continue
        except asyncio.CancelledError:
            break


# ------------------------------------------------------------------------------- in-process
class InProcessContext(NodeContext):
    """All nodes in one Python process: a class-level registry of asyncio queues.

    Payloads are handed over by reference -- a CUDA tensor sent between two device-resident nodes is not copied.
    Sending to an id that never started raises ``ValueError``.  The cheapest context; the default of
    :class:`~byzpy_b200.engine.peer_to_peer.train.PeerToPeer` here.

    Examples
    --------
    >>> import asyncio
    >>> from byzpy_b200.engine.graph.pool import ActorPoolConfig
    >>> from byzpy_b200.engine.node.application import NodeApplication
    >>> from byzpy_b200.engine.node.context import InProcessContext
    >>> from byzpy_b200.engine.node.decentralized import DecentralizedNode
    >>> async def demo():
    ...     nodes = [DecentralizedNode(node_id=i, application=NodeApplication(name=i, actor_pool=[ActorPoolConfig("thread")]),
    ...                                context=InProcessContext()) for i in ("a", "b")]
    ...     got = asyncio.get_running_loop().create_future()
    ...     async def on_ping(sender, payload):
    ...         got.set_result((sender, payload))
    ...     nodes[1].register_message_handler("ping", on_ping)
    ...     for n in nodes:
    ...         await n.start()
    ...     try:
    ...         await nodes[0].send_message("b", "ping", {"x": 1})
    ...         return await asyncio.wait_for(got, 5)
    ...     finally:
    ...         for n in nodes:
    ...             await n.shutdown()
    >>> asyncio.run(demo())
    ('a', {'x': 1})
    """

    _registry: Dict[Any, "InProcessContext"] = {}

    def __init__(self) -> None:
        self._node: Optional["DecentralizedNode"] = None
        self._inbox: asyncio.Queue = asyncio.Queue()
        self._running = False

    async def start(self, node: "DecentralizedNode") -> None:
        if self._running:
            return
        self._node = node
        self._running = True
        InProcessContext._registry[node.node_id] = self

    async def send_message(self, to_node_id: Any, message_type: str, payload: Any) -> None:
        if not self._running:
            raise RuntimeError("InProcessContext is not started.")
        target = InProcessContext._registry.get(to_node_id)
        if target is None or not target._running:
            raise ValueError(f"Target node {to_node_id} not found or not running.")
        await target._inbox.put({"from": self._node.node_id if self._node else "unknown",
                                 "type": message_type, "payload": payload})

    async def receive_messages(self) -> AsyncIterator[Any]:
        async for msg in _drain(self._inbox, lambda: self._running):
            yield msg

    async def shutdown(self) -> None:
        if not self._running:
            return
        self._running = False
        if self._node is not None and InProcessContext._registry.get(self._node.node_id) is self:
            del InProcessContext._registry[self._node.node_id]
        self._node = None
        while not self._inbox.empty():
            self._inbox.get_nowait()


# ---------------------------------------------------------------------------------- process
class _PipeReader(threading.Thread):
    """Blocks on ``conn.recv_bytes()`` and posts decoded messages onto an asyncio queue."""

    def __init__(self, conn, loop: asyncio.AbstractEventLoop, queue: asyncio.Queue) -> None:
        super().__init__(daemon=True)
        self.conn, self.loop, self.queue = conn, loop, queue

    def run(self) -> None:
        while True:
            try:
                blob = self.conn.recv_bytes()
            except (EOFError, OSError):
                break
            try:
                msg = cloudpickle.loads(blob)
            except Exception:
                continue
            try:
                self.loop.call_soon_threadsafe(self.queue.put_nowait, msg)
            except RuntimeError:
                break


class ProcessContext(NodeContext):
    """One OS process per node: the node is mirrored inside a spawned child and runs there.

    The child builds its own :class:`~byzpy_b200.engine.node.decentralized.DecentralizedNode` around the pickled
    application (an ``init_callback`` registered on the node runs in the child, so models and data loaders can be
    created where they are used), executes pipelines and autonomous tasks there, and talks to the parent over a duplex
    pipe; the parent relays child-to-child messages.  Reader threads wake the event loops directly -- no polling.

    Parameters
    ----------
    start_timeout : float, default 60.0
        Seconds to wait for the child to report ready (a fresh interpreter importing torch takes a few).

    Notes
    -----
    The reference's default context; here it is the default of ``DecentralizedCluster.add_node`` and selectable for
    ``PeerToPeer`` with ``context_factory=lambda node_id, index: ProcessContext()`` or ``BYZPY_P2P_CONTEXT=process``.
    """

    _registry: Dict[Any, "ProcessContext"] = {}

    def __init__(self, *, start_timeout: float = 60.0) -> None:
        self._node_id: Any = None
        self._process = None
        self._conn = None
        self._queue: Optional[asyncio.Queue] = None
        self._reader: Optional[_PipeReader] = None
        self._running = False
        self._start_timeout = start_timeout
        self._send_lock = threading.Lock()

    def _put(self, msg: dict) -> None:
        blob = cloudpickle.dumps(msg)
        with self._send_lock:
            self._conn.send_bytes(blob)

    async def start(self, node: "DecentralizedNode") -> None:
        if self._running:
            return
        self._node_id = node.node_id
        meta = {k: v for k, v in dict(getattr(node.scheduler, "metadata", {})).items() if k != "scheduler"}
        config = {"node_id": node.node_id, "application": node.application, "topology": node.topology,
                  "metadata": meta, "node_id_map": getattr(node, "_node_id_map", None)}
        for attr, key in (("_p2p_node_objects", "_node_objects"), ("_init_callback", "init_callback")):
            value = getattr(node, attr, None)
            if value is not None:
                config[key] = value
        if "_node_objects" in config:
            # The worker objects behind the pipelines are mirrored into the child so autonomous tasks can run
            # there; that is an optimisation, not a requirement (driver-paced pipelines execute in this
            # process).  User nodes holding unpicklable state (a live DataLoader iterator, ...) stay here only.
            try:
                cloudpickle.dumps(config["_node_objects"])
            except Exception:  # noqa: BLE001  (whatever the object's __getstate__ raises)
                del config["_node_objects"]
        try:
            blob = cloudpickle.dumps(config)
        except Exception as exc:  # noqa: BLE001
            bad = []
            for key, value in config.items():
                try:
                    cloudpickle.dumps(value)
                except Exception:  # noqa: BLE001
                    bad.append(key)
            raise RuntimeError(f"ProcessContext cannot ship node {node.node_id!r} to its child process: "
                               f"config entries {bad} are not picklable ({exc!r})") from exc
        ctx = process_context()      # fork-server with torch preloaded ("spawn" semantics, fast start)
        self._conn, child = ctx.Pipe(duplex=True)
        self._process = ctx.Process(target=_process_node_main, args=(blob, child), daemon=True)
        self._process.start()
        child.close()
        loop = asyncio.get_running_loop()
        self._queue = asyncio.Queue()
        self._reader = _PipeReader(self._conn, loop, self._queue)
        self._reader.start()
        self._running = True
        ProcessContext._registry[node.node_id] = self
        # wait for the child's readiness signal instead of sleeping a fixed amount
        deadline = loop.time() + self._start_timeout
        backlog = []
        while True:
            remaining = deadline - loop.time()
            if remaining <= 0 or not self._process.is_alive():
                raise RuntimeError(f"ProcessContext child for node {node.node_id!r} failed to start")
            try:
                msg = await asyncio.wait_for(self._queue.get(), timeout=min(remaining, 0.5))
            except asyncio.TimeoutError:
                continue
            if msg.get("type") == "_ready":
                break
            if msg.get("type") == "_fatal":
                raise RuntimeError(f"node process failed: {msg.get('payload')}")
            backlog.append(msg)
        for msg in backlog:
            self._queue.put_nowait(msg)
        ProcessContext._rebalance_threads()

    @classmethod
    def _rebalance_threads(cls) -> None:
        """Tell every running node process its share of the host's cores (``cores // node processes``): k
        children each opening full-width OpenMP regions spend their time spinning at barriers."""
        import os

        if os.environ.get("BYZPY_INTRAOP_GOVERNOR", "1") in ("0", "false", "False"):
            return
        running = [c for c in cls._registry.values() if c._running and c._conn is not None]
        share = max(1, (os.cpu_count() or 1) // max(1, len(running)))
        for ctx in running:
            try:
                ctx._put({"_command": "threads", "value": share})
            except Exception:
                pass

    async def send_message(self, to_node_id: Any, message_type: str, payload: Any) -> None:
        if not self._running:
            raise RuntimeError("ProcessContext is not started.")
        target = ProcessContext._registry.get(to_node_id)
        if target is None or not target._running:
            raise ValueError(f"Target node {to_node_id} not found or not running.")
        target._put({"from": self._node_id, "type": message_type, "payload": payload})

    async def receive_messages(self) -> AsyncIterator[Any]:
        if not self._running:
            raise RuntimeError("ProcessContext is not started.")
        async for msg in _drain(self._queue, lambda: self._running):
            if msg.get("_route_request"):
                target = ProcessContext._registry.get(msg.get("to"))
                if target is not None and target._running:
                    try:
                        target._put({"from": msg["from"], "type": msg["type"], "payload": msg["payload"]})
                    except Exception:
                        pass
                continue
            if msg.get("type") == "_message_received":
                yield {"from": msg["from"], "type": msg["message_type"], "payload": msg["payload"]}
                continue
            if msg.get("type") in ("_ready", "_fatal"):
                continue
            yield msg

    async def call_child(self, command: str, payload: Any = None) -> None:
        """Send a control command (e.g. ``"stop"``) to the mirror node."""
        self._put({"_command": command, "payload": payload})

    async def shutdown(self) -> None:
        if not self._running:
            return
        self._running = False
        try:
            self._put({"_command": "stop"})
        except Exception:
            pass
        proc = self._process
        if proc is not None:
            loop = asyncio.get_running_loop()
            await loop.run_in_executor(None, proc.join, 3.0)
            if proc.is_alive():
                proc.terminate()
                await loop.run_in_executor(None, proc.join, 1.0)
        if ProcessContext._registry.get(self._node_id) is self:
            del ProcessContext._registry[self._node_id]
            ProcessContext._rebalance_threads()
        try:
            self._conn.close()
        except Exception:
            pass
        self._conn = self._process = self._queue = None


class _SubprocessBridgeContext(NodeContext):
    """Context of the mirror node inside the child: everything goes through the parent's pipe."""

    def __init__(self, node_id: Any, conn, queue: asyncio.Queue) -> None:
        self._node_id = node_id
        self._conn = conn
        self._queue = queue
        self._running = False
        self._lock = threading.Lock()

    def _put(self, msg: dict) -> None:
        blob = cloudpickle.dumps(msg)
        with self._lock:
            self._conn.send_bytes(blob)

    async def start(self, node: "DecentralizedNode") -> None:
        self._running = True

    async def send_message(self, to_node_id: Any, message_type: str, payload: Any) -> None:
        if not self._running:
            raise RuntimeError("bridge context is not started")
        self._put({"_route_request": True, "to": to_node_id, "from": self._node_id,
                   "type": message_type, "payload": payload})

    async def receive_messages(self) -> AsyncIterator[Any]:
        async for msg in _drain(self._queue, lambda: self._running):
            if msg.get("_command") == "stop":
                self._running = False
                break
            if msg.get("_command") == "threads":       # this host's node processes split the cores
                try:
                    import torch

                    torch.set_num_threads(max(1, int(msg.get("value") or 1)))
                except Exception:
                    pass
                continue
            if "_command" in msg:
                continue
            # tell the parent-side node object about the delivery, then hand it to the mirror
            try:
                self._put({"type": "_message_received", "from": msg.get("from"),
                           "message_type": msg.get("type"), "payload": msg.get("payload")})
            except Exception:
                pass
            yield msg

    async def shutdown(self) -> None:
        self._running = False


def _process_node_main(config_blob: bytes, conn) -> None:
    """Entry point of a ``ProcessContext`` child: build the mirror node and serve messages."""
    import inspect

    async def run() -> None:
        from .decentralized import DecentralizedNode

        cfg = cloudpickle.loads(config_blob)
        loop = asyncio.get_running_loop()
        queue: asyncio.Queue = asyncio.Queue()
        _PipeReader(conn, loop, queue).start()
        bridge = _SubprocessBridgeContext(cfg["node_id"], conn, queue)
        node = DecentralizedNode(node_id=cfg["node_id"], application=cfg["application"],
                                 context=bridge, topology=cfg.get("topology"),
                                 metadata=cfg.get("metadata"), node_id_map=cfg.get("node_id_map"))
        if "_node_objects" in cfg:
            node._p2p_node_objects = cfg["_node_objects"]
            try:
                from ..peer_to_peer import runner as _runner

                _runner._NODE_OBJECT_REGISTRY.update(cfg["_node_objects"])
            except Exception:
                pass
        await node.start()
        callback = cfg.get("init_callback")
        if callback is not None:
            out = callback(node)
            if inspect.isawaitable(out):
                await out
        bridge._put({"type": "_ready", "from": cfg["node_id"]})
        while bridge._running:
            await asyncio.sleep(0.05)
        await node.shutdown()

    try:
        asyncio.run(run())
    except BaseException as exc:  # noqa: BLE001
        try:
            conn.send_bytes(cloudpickle.dumps({"type": "_fatal", "payload": repr(exc)}))
        except Exception:
            pass
    finally:
        try:
            conn.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------ remote
class RemoteContext(NodeContext):
    """Client of a :class:`~byzpy_b200.engine.node.remote_server.RemoteNodeServer` hub: every message goes to the
    hub, which forwards it to the addressee's connection.

    Parameters
    ----------
    host, port :
        Address of the hub.
    gpu_direct : bool, default False
        Ship CUDA tensors as CUDA-IPC handles (hub and nodes on one machine) instead of staging them through the host.
    """

    def __init__(self, host: str, port: int, *, gpu_direct: bool = False):
        self.host, self.port = host, int(port)
        self.gpu_direct = gpu_direct
        self._client = None
        self._node: Optional["DecentralizedNode"] = None
        self._running = False

    async def start(self, node: "DecentralizedNode") -> None:
        if self._running:
            return
        from .remote_client import RemoteNodeClient

        self._node = node
        self._client = RemoteNodeClient(self.host, self.port, gpu_direct=self.gpu_direct)
        try:
            await self._client.connect(timeout=5.0)
        except (ConnectionError, OSError, asyncio.TimeoutError) as exc:
            raise ConnectionError(f"Failed to connect to {self.host}:{self.port}") from exc
        self._running = True
        await self._client.register_node(node.node_id)

    async def send_message(self, to_node_id: Any, message_type: str, payload: Any) -> None:
        if not self._running or self._client is None:
            raise RuntimeError("RemoteContext is not started.")
        await self._client.send_message(to_node_id, message_type, payload,
                                        from_node_id=self._node.node_id if self._node else None)

    async def receive_messages(self) -> AsyncIterator[Any]:
        while self._running:
            try:
                msg = await self._client.receive_message(timeout=0.1)
            except asyncio.CancelledError:
                break
            if msg is None:
                if not self._client.is_connected():
                    await asyncio.sleep(0.05)
                continue
            yield msg

    async def shutdown(self) -> None:
        if not self._running:
            return
        self._running = False
        if self._client is not None:
            await self._client.disconnect()
            self._client = None
        self._node = None


# -------------------------------------------------------------------------------------- mesh
class MeshRemoteContext(NodeContext):
    """Full mesh over TCP: every node listens on its own address and dials every peer.

    Parameters
    ----------
    local_host, local_port :
        Where this node listens.
    peer_addresses : dict
        ``{node id: (host, port)}`` of the other nodes.
    connect_timeout : float, default 5.0
    reconnect_interval : float, default 2.0
        Peers that are down are re-dialled this often in the background; nodes may start in any order.
    gpu_direct : bool, default False
        Ship CUDA tensors as CUDA-IPC handles between processes of one machine.

    Notes
    -----
    A send uses the outbound connection to the peer and falls back to the connection the peer opened to us; when
    neither exists it raises :class:`PeerUnreachableError`.
    """

    def __init__(self, local_host: str, local_port: int, peer_addresses: Dict[Any, Tuple[str, int]],
                 connect_timeout: float = 5.0, reconnect_interval: float = 2.0, *,
                 gpu_direct: bool = False):
        self.local_host, self.local_port = local_host, int(local_port)
        self.peer_addresses = dict(peer_addresses)
        self.connect_timeout = connect_timeout
        self.reconnect_interval = reconnect_interval
        self.gpu_direct = gpu_direct
        self._node: Optional["DecentralizedNode"] = None
        self._running = False
        self._local_server: Optional[asyncio.AbstractServer] = None
        self._peer_clients: Dict[Any, Any] = {}
        self._inbound_connections: Dict[Any, asyncio.StreamWriter] = {}
        self._inbox: asyncio.Queue = asyncio.Queue()
        self._monitor_task: Optional[asyncio.Task] = None
        self._serve_task: Optional[asyncio.Task] = None

    async def start(self, node: "DecentralizedNode") -> None:
        if self._running:
            return
        self._node = node
        self._running = True
        self._local_server = await asyncio.start_server(self._handle_inbound_connection,
                                                        self.local_host, self.local_port)
        self.local_port = self._local_server.sockets[0].getsockname()[1]
        self._serve_task = asyncio.ensure_future(self._local_server.serve_forever())
        await self._dial_missing()
        self._monitor_task = asyncio.ensure_future(self._connection_monitor())

    async def _dial(self, peer_id: Any) -> bool:
        from .remote_client import RemoteNodeClient

        host, port = self.peer_addresses[peer_id]
        client = RemoteNodeClient(host, port, gpu_direct=self.gpu_direct)
        try:
            await client.connect(timeout=self.connect_timeout)
            await client.register_node(self._node.node_id)
        except Exception:
            try:
                await client.disconnect()
            except Exception:
                pass
            return False
        self._peer_clients[peer_id] = client
        return True

    async def _dial_missing(self) -> None:
        for peer_id in self.peer_addresses:
            if self._node is not None and peer_id == self._node.node_id:
                continue
            client = self._peer_clients.get(peer_id)
            if client is not None and client.is_connected():
                continue
            if client is not None:
                try:
                    await client.disconnect()
                except Exception:
                    pass
                self._peer_clients.pop(peer_id, None)
            await self._dial(peer_id)

    async def _connection_monitor(self) -> None:
        try:
            while self._running:
                await asyncio.sleep(self.reconnect_interval)
                if self._running:
                    await self._dial_missing()
        except asyncio.CancelledError:
            pass

    async def _handle_inbound_connection(self, reader: asyncio.StreamReader,
                                         writer: asyncio.StreamWriter) -> None:
        from .remote_client import read_frame

        peer_id = None
        try:
            while self._running:
                try:
                    msg = await read_frame(reader)
                except (asyncio.IncompleteReadError, ConnectionError, OSError):
                    break
                if msg.get("type") == "_register_node":
                    peer_id = msg.get("node_id")
                    self._inbound_connections[peer_id] = writer
                    continue
                await self._inbox.put({"from": msg.get("from", peer_id if peer_id is not None else "unknown"),
                                       "type": msg.get("type", "unknown"), "payload": msg.get("payload")})
        finally:
            if peer_id is not None and self._inbound_connections.get(peer_id) is writer:
                self._inbound_connections.pop(peer_id, None)
            try:
                writer.close()
            except Exception:
                pass

    async def send_message(self, to_node_id: Any, message_type: str, payload: Any) -> None:
        from .remote_client import write_frame

        if not self._running:
            raise RuntimeError("MeshRemoteContext is not started.")
        sender = self._node.node_id if self._node else None
        client = self._peer_clients.get(to_node_id)
        if client is not None and client.is_connected():
            try:
                await client.send_message(to_node_id, message_type, payload, from_node_id=sender)
                return
            except RuntimeError:
                pass  # fall back to the inbound connection
        writer = self._inbound_connections.get(to_node_id)
        if writer is not None and not writer.is_closing():
            await write_frame(writer, {"to": to_node_id, "from": sender, "type": message_type,
                                       "payload": payload}, gpu_direct=self.gpu_direct)
            return
        if to_node_id in self.peer_addresses and await self._dial(to_node_id):
            await self._peer_clients[to_node_id].send_message(to_node_id, message_type, payload,
                                                              from_node_id=sender)
            return
        raise PeerUnreachableError(f"No connection to peer {to_node_id!r}")

    async def receive_messages(self) -> AsyncIterator[Any]:
        # outbound clients may also carry replies written on the same socket by the peer
        async def pump(client) -> None:
            while self._running and client.is_connected():
                msg = await client.receive_message(timeout=0.2)
                if msg is not None:
                    await self._inbox.put({"from": msg.get("from", "unknown"),
                                           "type": msg.get("type", "unknown"),
                                           "payload": msg.get("payload")})

        pumps: Dict[Any, asyncio.Task] = {}
        try:
            while self._running:
                for pid, client in list(self._peer_clients.items()):
                    t = pumps.get(pid)
                    if (t is None or t.done()) and client.is_connected():
                        pumps[pid] = asyncio.ensure_future(pump(client))
                try:
                    yield await asyncio.wait_for(self._inbox.get(), timeout=0.1)
                except asyncio.TimeoutError:
                    continue
                except asyncio.CancelledError:
                    break
        finally:
            for t in pumps.values():
                t.cancel()

    async def shutdown(self) -> None:
        if not self._running:
            return
        self._running = False
        if self._monitor_task is not None:
            self._monitor_task.cancel()
            try:
                await self._monitor_task
            except (asyncio.CancelledError, Exception):
                pass
        for client in list(self._peer_clients.values()):
            try:
                await client.disconnect()
            except Exception:
                pass
        self._peer_clients.clear()
        for w in list(self._inbound_connections.values()):
            try:
                w.close()
            except Exception:
                pass
        self._inbound_connections.clear()
        if self._serve_task is not None:
            self._serve_task.cancel()
            try:
                await self._serve_task
            except (asyncio.CancelledError, Exception):
                pass
            self._serve_task = None
        if self._local_server is not None:
            self._local_server.close()
            try:
                await asyncio.wait_for(self._local_server.wait_closed(), timeout=2.0)
            except Exception:
                pass
            self._local_server = None
        self._node = None

    def get_connected_peers(self) -> list:
        out = [pid for pid, c in self._peer_clients.items() if c.is_connected()]
        out += [pid for pid, w in self._inbound_connections.items() if not w.is_closing() and pid not in out]
        return out


__all__ = ["NodeContext", "InProcessContext", "ProcessContext", "RemoteContext", "MeshRemoteContext",
           "PeerUnreachableError"]
