"""Transport backed by one :class:`TcpMailbox` + poller thread per registered node
(reference engine/transport/tcp.py:15-57)."""
from __future__ import annotations

import queue
import threading
from typing import Any, Callable, Dict

from .tcp_simple import TcpMailbox, send_message


class TcpTransport:
    def __init__(self) -> None:
        self._mailboxes: Dict[str, TcpMailbox] = {}
        self._threads: Dict[str, threading.Thread] = {}
        self._stop = threading.Event()

    def register(self, node_id: str, handler: Callable[[Any], None]) -> None:
        if node_id in self._mailboxes:
            raise ValueError(f"Node {node_id} already registered")
        box = TcpMailbox()
        self._mailboxes[node_id] = box

        def poll() -> None:
            while not self._stop.is_set():
                try:
                    msg = box.recv(timeout=0.1)
                except queue.Empty:
                    continue
                try:
                    handler(msg)
                except Exception:
                    continue

        t = threading.Thread(target=poll, daemon=True)
        self._threads[node_id] = t
        t.start()

    def send(self, to_id: str, payload: Any) -> None:
        box = self._mailboxes.get(to_id)
        if box is None:
            raise KeyError(f"Unknown node_id {to_id}")
        send_message(("127.0.0.1", box.port), payload)

    def close(self) -> None:
        self._stop.set()
        for box in self._mailboxes.values():
            box.close()
        for t in self._threads.values():
            t.join(timeout=1.0)


__all__ = ["TcpTransport"]
