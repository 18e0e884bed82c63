"""One-shot TCP mailbox for the legacy runners: a sender opens a connection, writes ONE frame
(8-byte big-endian length + pickle body -- the wire layout of reference
engine/transport/tcp_simple.py:26-80) and closes; the mailbox queues decoded payloads.

Built on ``socketserver.ThreadingTCPServer`` (one daemon thread per connection); malformed frames
are dropped, never raised into the accept loop.
"""
from __future__ import annotations

import pickle
import queue
import socket
import socketserver
import struct
import threading
from typing import Any, Optional, Tuple

_LEN = struct.Struct(">Q")


def _recv_exactly(sock: socket.socket, nbytes: int) -> bytes:
    buf = bytearray()
    while len(buf) < nbytes:
        part = sock.recv(min(1 << 20, nbytes - len(buf)))
        if not part:
            raise ConnectionError("connection closed before receiving payload")
        buf += part
    return bytes(buf)


def send_message(addr: Tuple[str, int], payload: Any) -> None:
    """Deliver ``payload`` to the mailbox listening at ``addr`` (blocking, 5 s connect timeout)."""
    blob = pickle.dumps(payload, protocol=pickle.HIGHEST_PROTOCOL)
    with socket.create_connection(addr, timeout=5.0) as sock:
        sock.sendall(_LEN.pack(len(blob)))
        sock.sendall(blob)


class _FrameHandler(socketserver.BaseRequestHandler):
    def handle(self) -> None:
        try:
            (size,) = _LEN.unpack(_recv_exactly(self.request, _LEN.size))
            self.server.inbox.put(pickle.loads(_recv_exactly(self.request, size)))  # type: ignore[attr-defined]
        except Exception:
            return


class _Server(socketserver.ThreadingTCPServer):
    allow_reuse_address = True
    daemon_threads = True


class TcpMailbox:
    """A listening socket with a queue behind it: the smallest possible TCP inbox.

    Each accepted connection delivers one length-prefixed pickled payload.  ``recv(timeout=None)`` returns the next one
    (``queue.Empty`` on timeout), ``host`` / ``port`` say where to send (``port=0`` picks a free one), ``close()`` stops the
    server thread.  Payloads are unpickled: loopback or trusted networks only.

    Examples
    --------
    >>> from byzpy_b200.engine.transport.tcp_simple import TcpMailbox, send_message
    >>> box = TcpMailbox()
    >>> send_message((box.host, box.port), {"hello": 1})
    >>> box.recv(timeout=5)
    {'hello': 1}
    >>> box.close()
    """

    def __init__(self, host: str = "127.0.0.1", port: int = 0) -> None:
        self._srv = _Server((host, port), _FrameHandler)
        self._srv.inbox = queue.Queue()  # type: ignore[attr-defined]
        self.host, self.port = host, self._srv.server_address[1]
        self._thread = threading.Thread(target=self._srv.serve_forever, kwargs={"poll_interval": 0.1}, daemon=True)
        self._thread.start()

    def recv(self, timeout: Optional[float] = None) -> Any:
        """Next payload; raises ``queue.Empty`` after ``timeout`` seconds."""
        return self._srv.inbox.get(timeout=timeout)  # type: ignore[attr-defined]

    def close(self) -> None:
        self._srv.shutdown()
        self._srv.server_close()
        self._thread.join(timeout=1.0)


__all__ = ["TcpMailbox", "send_message"]
