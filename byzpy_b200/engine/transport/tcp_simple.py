"""Loopback-friendly TCP mailbox: one connection per message, 8-byte big-endian length +
pickle body (frame layout of reference engine/transport/tcp_simple.py:26-80)."""
from __future__ import annotations

import pickle
import queue
import socket
import struct
import threading
from typing import Any, Optional, Tuple

_HDR = struct.Struct(">Q")


def _read_exact(conn: socket.socket, n: int) -> bytes:
    chunks, got = [], 0
    while got < n:
        part = conn.recv(min(1 << 20, n - got))
        if not part:
            raise ConnectionError("connection closed before receiving payload")
        chunks.append(part)
        got += len(part)
    return b"".join(chunks)


def send_message(addr: Tuple[str, int], payload: Any) -> None:
    body = pickle.dumps(payload, protocol=pickle.HIGHEST_PROTOCOL)
    with socket.create_connection(addr, timeout=5.0) as sock:
        sock.sendall(_HDR.pack(len(body)) + body)


class TcpMailbox:
    def __init__(self, host: str = "127.0.0.1", port: int = 0) -> None:
        self.host = host
        self._q: "queue.Queue[Any]" = queue.Queue()
        self._stop = threading.Event()
        self._server = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self._server.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self._server.bind((host, port))
        self._server.listen()
        self._server.settimeout(0.2)
        self.port = self._server.getsockname()[1]
        self._thread = threading.Thread(target=self._accept_loop, daemon=True)
        self._thread.start()

    def _accept_loop(self) -> None:
        while not self._stop.is_set():
            try:
                conn, _ = self._server.accept()
            except socket.timeout:
                continue
            except OSError:
                break
            threading.Thread(target=self._serve, args=(conn,), daemon=True).start()

    def _serve(self, conn: socket.socket) -> None:
        with conn:
            try:
                (n,) = _HDR.unpack(_read_exact(conn, _HDR.size))
                self._q.put(pickle.loads(_read_exact(conn, n)))
            except Exception:
                return  # malformed frame: drop

    def recv(self, timeout: Optional[float] = None) -> Any:
        return self._q.get(timeout=timeout)

    def close(self) -> None:
        self._stop.set()
        try:
            self._server.close()
        except Exception:
            pass
        self._thread.join(timeout=1.0)


__all__ = ["TcpMailbox", "send_message"]
