"""``NodeCluster``: a registry of named :class:`NodeRunner` processes with a pluggable message
transport (legacy building block, reference engine/node_cluster.py:16-60).

Without a transport, ``send`` posts straight into the runner's inbox pipe; with one
(``LocalTransport`` / ``TcpTransport``) every node registers its inbox as the delivery callback and
``send`` goes through the transport, so the same script works in-process or over TCP.
"""
from __future__ import annotations

import time
from typing import Any, Callable, Dict, Iterator, Optional

from .node_runner import NodeRunner

StepFn = Callable[[dict], dict]
MessageFn = Callable[[dict, Any], dict]


class NodeCluster:
    """Named :class:`~byzpy_b200.engine.node_runner.NodeRunner` processes plus an optional message transport (legacy).

    ``add_node(node_id, step_fn, msg_handler, init_state=None)``, ``start_all()``, ``stop_all()``, ``send(node_id, msg)``,
    ``state(node_id)``.  With ``transport=LocalTransport()`` or ``TcpTransport()`` messages go through the transport (the
    same script then runs in-process or over loopback sockets); without one they are written to the runner's pipe.
    ``examples/p2p/decentralized_demo.py`` is a complete program.
    """

    def __init__(self, transport=None) -> None:
        self._transport = transport
        self._nodes: Dict[str, NodeRunner] = {}

    # -- membership -------------------------------------------------------------------------
    def add_node(self, node_id: str, step_fn: StepFn, msg_handler: MessageFn, *,
                 init_state: Optional[dict] = None) -> None:
        if node_id in self._nodes:
            raise ValueError(f"Node {node_id} already exists")
        self._nodes[node_id] = node = NodeRunner(step_fn, msg_handler, init_state=init_state)
        if self._transport is not None:
            self._transport.register(node_id, node.send_message)

    def _runner(self, node_id: str) -> NodeRunner:
        try:
            return self._nodes[node_id]
        except KeyError:
            raise KeyError(f"unknown node {node_id!r}") from None

    def __iter__(self) -> Iterator[str]:
        return iter(self._nodes)

    def __len__(self) -> int:
        return len(self._nodes)

    # -- lifecycle ----------------------------------------------------------------------------
    def start_all(self) -> None:
        for runner in self._nodes.values():
            runner.start()

    def stop_all(self) -> None:
        for runner in self._nodes.values():
            runner.stop()

    def start_auto(self, node_id: str, interval_sec: float) -> None:
        self._runner(node_id).start_auto(interval_sec)

    # -- messaging / inspection ---------------------------------------------------------------
    def send(self, to_id: str, msg: Any) -> None:
        if self._transport is None:
            self._runner(to_id).send_message(msg)
        else:
            self._transport.send(to_id, msg)

    def state(self, node_id: str) -> dict:
        return self._runner(node_id).state()

    def barrier(self, duration: float) -> None:
        """Crude time barrier kept for API parity (lets asynchronous transports land messages)."""
        time.sleep(max(0.0, float(duration)))


__all__ = ["NodeCluster"]
