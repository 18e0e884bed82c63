"""``TcpTransport``: every registered node gets a loopback :class:`TcpMailbox` and a pump thread that
feeds received payloads to the node's handler (reference engine/transport/tcp.py:15-57).  Messages
therefore really cross a socket even on one host, which is what the legacy runner tests rely on.
"""
from __future__ import annotations

import queue
import threading
from typing import Any, Callable, Dict, NamedTuple

from .tcp_simple import TcpMailbox, send_message


class _Route(NamedTuple):
    mailbox: TcpMailbox
    pump: threading.Thread


class TcpTransport:
    """Loopback-socket :class:`~byzpy_b200.engine.transport.base.Transport`: every registered node gets a
    :class:`~byzpy_b200.engine.transport.tcp_simple.TcpMailbox` on a free port and a pump thread that feeds its handler;
    ``send`` connects to the addressee's port and writes one length-prefixed pickle.  ``close()`` stops the pumps and
    the listeners.

    Parameters
    ----------
    host : str, default "127.0.0.1"
    """

    def __init__(self, host: str = "127.0.0.1") -> None:
        self._host = host
        self._routes: Dict[str, _Route] = {}
        self._closing = threading.Event()

    def _pump(self, mailbox: TcpMailbox, handler: Callable[[Any], None]) -> None:
        while not self._closing.is_set():
            try:
                payload = mailbox.recv(timeout=0.1)
            except queue.Empty:
                continue
            try:
                handler(payload)
            except Exception:       # a failing handler must not kill delivery to the node
                pass

    def register(self, node_id: str, handler: Callable[[Any], None]) -> None:
        if node_id in self._routes:
            raise ValueError(f"Node {node_id} already registered")
        mailbox = TcpMailbox(self._host)
        pump = threading.Thread(target=self._pump, args=(mailbox, handler), daemon=True,
                                name=f"tcp-transport-{node_id}")
        self._routes[node_id] = _Route(mailbox, pump)
        pump.start()

    def address_of(self, node_id: str):
        route = self._routes[node_id]
        return route.mailbox.host, route.mailbox.port

    def send(self, to_id: str, payload: Any) -> None:
        route = self._routes.get(to_id)
        if route is None:
            raise KeyError(f"Unknown node_id {to_id}")
        send_message((self._host, route.mailbox.port), payload)

    def close(self) -> None:
        self._closing.set()
        for route in self._routes.values():
            route.mailbox.close()
        for route in self._routes.values():
            route.pump.join(timeout=1.0)


__all__ = ["TcpTransport"]
