"""TCP client side of the decentralized node runtime
(reference engine/node/remote_client.py:11-278).

Framing: 4-byte big-endian length + cloudpickle body.  ``serialize_message`` moves CUDA tensors
to the host by default; with ``gpu_direct=True`` (same NVSwitch box) CUDA tensors travel as CUDA
IPC handles and arrive zero-copy on the peer's side.
"""
from __future__ import annotations

import asyncio
import struct
from typing import Any, Dict, Optional

import cloudpickle

from ..actor.transports import cuda_ipc

_LEN = struct.Struct(">I")


def serialize_message(msg: Dict[str, Any], *, gpu_direct: bool = False) -> bytes:
    """Message dict -> bytes: ``b"G"`` + CUDA-IPC pickle when ``gpu_direct`` (tensors stay on the device), else ``b"P"`` +
    cloudpickle with tensors moved to the host.
    """
    if gpu_direct and cuda_ipc.available():
        return b"G" + cuda_ipc.dumps(msg, same_host=True)
    return b"P" + cloudpickle.dumps(cuda_ipc._to_host(msg))


def deserialize_message(data: bytes) -> Dict[str, Any]:
    """Inverse of :func:`serialize_message` (untagged frames are read as plain cloudpickle)."""
    tag, body = data[:1], data[1:]
    if tag == b"G":
        return cuda_ipc.loads(body)
    if tag == b"P":
        return cloudpickle.loads(body)
    return cloudpickle.loads(data)  # untagged legacy frame


async def write_frame(writer: asyncio.StreamWriter, msg: Dict[str, Any], *, gpu_direct: bool = False) -> None:
    """Write one length-prefixed message to the stream and drain it."""
    body = serialize_message(msg, gpu_direct=gpu_direct)
    writer.write(_LEN.pack(len(body)) + body)
    await writer.drain()


async def read_frame(reader: asyncio.StreamReader) -> Dict[str, Any]:
    """Read one length-prefixed message from the stream."""
    (n,) = _LEN.unpack(await reader.readexactly(_LEN.size))
    return deserialize_message(await reader.readexactly(n))


class RemoteNodeClient:
    """One TCP connection to a :class:`~byzpy_b200.engine.node.remote_server.RemoteNodeServer`.

    ``await connect(timeout=5.0)``; ``await register_node(node_id)`` announces an id this client owns;
    ``await send_message(to, type, payload)``; ``await receive_message(timeout=None)`` returns the next message the
    server forwarded (``None`` on timeout); ``is_connected()``; ``await disconnect()``.  Used by :class:`~byzpy_b200.engine.node.context.RemoteContext`; frames are written under a
    lock so concurrent senders cannot interleave.

    Parameters
    ----------
    host, port :
    gpu_direct : bool, default False
    """

    def __init__(self, host: str, port: int, *, gpu_direct: bool = False):
        self.host, self.port = host, int(port)
        self.gpu_direct = gpu_direct
        self._reader: Optional[asyncio.StreamReader] = None
        self._writer: Optional[asyncio.StreamWriter] = None
        self._connected = False
        self._running = False
        self._receive_task: Optional[asyncio.Task] = None
        self._message_queue: asyncio.Queue = asyncio.Queue()
        self._send_lock = asyncio.Lock()

    async def connect(self, timeout: float = 5.0) -> None:
        if self._connected:
            return
        try:
            self._reader, self._writer = await asyncio.wait_for(
                asyncio.open_connection(self.host, self.port), timeout=timeout)
        except asyncio.TimeoutError:
            raise asyncio.TimeoutError(f"Connection to {self.host}:{self.port} timed out") from None
        except OSError as exc:
            raise ConnectionError(f"Failed to connect to {self.host}:{self.port}") from exc
        self._connected = self._running = True
        self._receive_task = asyncio.ensure_future(self._receive_loop())

    async def disconnect(self) -> None:
        self._running = self._connected = False
        task, self._receive_task = self._receive_task, None
        if task is not None and not task.done():
            task.cancel()
            try:
                await task
            except (asyncio.CancelledError, Exception):
                pass
        if self._writer is not None:
            try:
                self._writer.close()
                await self._writer.wait_closed()
            except Exception:
                pass
        self._reader = self._writer = None

    def is_connected(self) -> bool:
        w = self._writer
        alive = (self._connected and w is not None and not w.is_closing()
                 and w.transport is not None and not w.transport.is_closing())
        if not alive:
            self._connected = False
        return alive

    async def _send(self, msg: Dict[str, Any]) -> None:
        if not self.is_connected():
            raise RuntimeError("Client is not connected")
        try:
            async with self._send_lock:
                await write_frame(self._writer, msg, gpu_direct=self.gpu_direct)
        except (BrokenPipeError, ConnectionResetError, OSError) as exc:
            self._connected = False
            raise RuntimeError(f"Connection lost: {exc}") from exc

    async def send_message(self, to_node_id: str, message_type: str, payload: Any,
                           from_node_id: Optional[str] = None) -> None:
        msg = {"to": to_node_id, "type": message_type, "payload": payload}
        if from_node_id:
            msg["from"] = from_node_id
        await self._send(msg)

    async def register_node(self, node_id: str) -> None:
        try:
            await self._send({"type": "_register_node", "node_id": node_id})
        except RuntimeError as exc:
            raise RuntimeError(f"Failed to register node: {exc}") from exc

    async def receive_message(self, timeout: Optional[float] = None) -> Optional[Dict[str, Any]]:
        try:
            if timeout is None:
                return await self._message_queue.get()
            return await asyncio.wait_for(self._message_queue.get(), timeout=timeout)
        except asyncio.TimeoutError:
            return None

    async def _receive_loop(self) -> None:
        try:
            while self._running:
                try:
                    msg = await read_frame(self._reader)
                except (asyncio.IncompleteReadError, ConnectionError, OSError):
                    self._connected = False
                    break
                await self._message_queue.put(msg)
        except asyncio.CancelledError:
            pass


__all__ = ["RemoteNodeClient", "serialize_message", "deserialize_message", "read_frame", "write_frame"]
