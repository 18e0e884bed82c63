"""Hub server hosting decentralized nodes and relaying between remote clients
(reference engine/node/remote_server.py:15-274).

Nodes registered on the server use ``ServerNodeContext``: local delivery goes through the
in-process registry, anything else is forwarded to the remote client that announced the target
id with a ``_register_node`` handshake.  Remote clients address any node id; the server delivers
locally or relays to the owning client.
"""
from __future__ import annotations

import asyncio
from typing import Any, Dict, Optional

from .context import InProcessContext, NodeContext
from .remote_client import read_frame, write_frame


class ServerNodeContext(NodeContext):
    """Context of a node hosted *on* a :class:`RemoteNodeServer`: peers registered on the same server are reached
    through the in-process registry, everybody else through the server's connection to the client that announced that
    id.
    """

    def __init__(self, server: "RemoteNodeServer", in_process_context: InProcessContext):
        self.server = server
        self.inner = in_process_context
        self._node = None

    async def start(self, node) -> None:
        self._node = node
        await self.inner.start(node)

    async def send_message(self, to_node_id, message_type: str, payload: Any) -> None:
        if to_node_id in InProcessContext._registry:
            await self.inner.send_message(to_node_id, message_type, payload)
            return
        await self.server.send_message_to_client(
            to_node_id, {"from": self._node.node_id if self._node else "unknown",
                         "type": message_type, "payload": payload})

    async def receive_messages(self):
        async for msg in self.inner.receive_messages():
            yield msg

    async def shutdown(self) -> None:
        await self.inner.shutdown()


class RemoteNodeServer:
    """Hub of the hub-and-spoke deployment: relays messages between
    :class:`~byzpy_b200.engine.node.context.RemoteContext` clients, and can host nodes itself.

    Parameters
    ----------
    host : str, default "localhost"
    port : int, default 8888
        0 picks a free port (``port`` holds it after ``start()``).
    gpu_direct : bool, default False
        Forward CUDA tensors as CUDA-IPC handles (all clients on this machine).

    Notes
    -----
    ``await start()`` binds, ``await serve()`` serves until cancelled, ``await shutdown()`` closes every connection,
    ``await register_node(node)`` hosts a :class:`~byzpy_b200.engine.node.decentralized.DecentralizedNode` on the
    server.  A client announces the node ids it owns with a ``_register_node`` frame; frames are length-prefixed
    pickles, so expose the port on trusted networks only.  ``examples/p2p/remote_tcp/server.py`` runs one.
    """

    def __init__(self, host: str = "localhost", port: int = 8888, *, gpu_direct: bool = False):
        self.host, self.port = host, int(port)
        self.gpu_direct = gpu_direct
        self._nodes: Dict[Any, Any] = {}
        self._clients: Dict[Any, asyncio.StreamWriter] = {}
        self._writers: set = set()          # every accepted connection, registered or not
        self._server: Optional[asyncio.AbstractServer] = None
        self._running = False

    async def register_node(self, node) -> None:
        if node.node_id in self._nodes:
            raise ValueError(f"Node {node.node_id!r} already registered")
        inner = node.context if isinstance(node.context, InProcessContext) else InProcessContext()
        node.context = ServerNodeContext(self, inner)
        self._nodes[node.node_id] = node
        await node.start()

    async def start(self) -> None:
        if self._server is None:
            self._server = await asyncio.start_server(self._handle_client, self.host, self.port)
            self.port = self._server.sockets[0].getsockname()[1]
            self._running = True

    async def serve(self) -> None:
        await self.start()
        try:
            await self._server.serve_forever()
        finally:
            # cancelled or stopped: also hang up on the clients that are still connected -- since Python
            # 3.12 ``Server.wait_closed()`` waits for every open connection, so leaving them would block
            self._running = False
            self._close_connections()
            self._server.close()

    def _close_connections(self) -> None:
        for w in list(self._writers):
            try:
                w.close()
            except Exception:
                pass
        self._writers.clear()

    async def _deliver(self, msg: Dict[str, Any]) -> None:
        target = msg.get("to")
        node = self._nodes.get(target)
        if node is not None:
            await node.handle_incoming_message(msg.get("from", "unknown"), msg.get("type", "unknown"),
                                               msg.get("payload"))
            return
        await self.send_message_to_client(target, {"from": msg.get("from", "unknown"),
                                                   "type": msg.get("type"), "payload": msg.get("payload")})

    async def _handle_client(self, reader: asyncio.StreamReader, writer: asyncio.StreamWriter) -> None:
        owned = []
        self._writers.add(writer)
        try:
            while True:
                try:
                    msg = await read_frame(reader)
                except (asyncio.IncompleteReadError, ConnectionError, OSError):
                    break
                if msg.get("type") == "_register_node":
                    nid = msg.get("node_id")
                    self._clients[nid] = writer
                    owned.append(nid)
                    continue
                try:
                    await self._deliver(msg)
                except Exception:
                    continue
        finally:
            for nid in owned:
                if self._clients.get(nid) is writer:
                    self._clients.pop(nid, None)
            self._writers.discard(writer)
            try:
                writer.close()
            except Exception:
                pass

    @property
    def _client_connections(self) -> Dict[Any, asyncio.StreamWriter]:
        """All open client connections: by node id once a client registered one, by connection identity
        before that (the reference's attribute name)."""
        named = {id(w) for w in self._clients.values()}
        out: Dict[Any, asyncio.StreamWriter] = dict(self._clients)
        out.update({f"conn-{id(w):x}": w for w in self._writers if id(w) not in named})
        return out

    async def send_message_to_client(self, client_writer, from_node_id=None, message_type: Optional[str] = None,
                                     payload: Any = None) -> None:
        """Two calling conventions: ``(node_id, message_dict)`` -- used inside this package, the registered
        client of that node id is looked up -- and the reference's ``(client_writer, from_node_id, message_type,
        payload)`` with the client's ``StreamWriter`` (reference engine/node/remote_server.py:226-250)."""
        if isinstance(client_writer, asyncio.StreamWriter) or hasattr(client_writer, "drain"):
            await write_frame(client_writer, {"from": from_node_id, "type": message_type, "payload": payload},
                              gpu_direct=self.gpu_direct)
            return
        node_id, msg = client_writer, from_node_id
        writer = self._clients.get(node_id)
        if writer is None or writer.is_closing():
            raise ValueError(f"Target node {node_id} not found on server or among clients")
        out = dict(msg)
        out.setdefault("to", node_id)
        await write_frame(writer, out, gpu_direct=self.gpu_direct)

    async def shutdown(self) -> None:
        self._running = False
        for node in list(self._nodes.values()):
            try:
                await node.shutdown()
            except Exception:
                pass
        self._nodes.clear()
        for w in list(self._clients.values()):
            try:
                w.close()
            except Exception:
                pass
        self._clients.clear()
        self._close_connections()
        if self._server is not None:
            self._server.close()
            try:
                await self._server.wait_closed()
            except Exception:
                pass
            self._server = None


__all__ = ["RemoteNodeServer", "ServerNodeContext"]
