"""A named set of :class:`NodeRunner` processes with an optional pluggable transport
(reference engine/node_cluster.py:16-60)."""
from __future__ import annotations

import time
from typing import Any, Callable, Dict, Optional

from .node_runner import NodeRunner


class NodeCluster:
    def __init__(self, transport=None) -> None:
        self._nodes: Dict[str, NodeRunner] = {}
        self._transport = transport

    def add_node(self, node_id: str, step_fn: Callable[[dict], dict],
                 msg_handler: Callable[[dict, Any], dict], *, init_state: Optional[dict] = None) -> None:
        if node_id in self._nodes:
            raise ValueError(f"Node {node_id} already exists")
        runner = NodeRunner(step_fn, msg_handler, init_state=init_state)
        self._nodes[node_id] = runner
        if self._transport is not None:
            self._transport.register(node_id, runner.send_message)

    def start_all(self) -> None:
        for n in self._nodes.values():
            n.start()

    def stop_all(self) -> None:
        for n in self._nodes.values():
            n.stop()

    def start_auto(self, node_id: str, interval_sec: float) -> None:
        self._nodes[node_id].start_auto(interval_sec)

    def send(self, to_id: str, msg: Any) -> None:
        if self._transport is not None:
            self._transport.send(to_id, msg)
        else:
            self._nodes[to_id].send_message(msg)

    def state(self, node_id: str) -> dict:
        return self._nodes[node_id].state()

    def barrier(self, duration: float) -> None:
        time.sleep(duration)


__all__ = ["NodeCluster"]
