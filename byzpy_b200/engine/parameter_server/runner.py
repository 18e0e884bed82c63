"""Prototype parameter server over ``NodeRunner`` processes (reference
engine/parameter_server/runner.py:49-90): worker runners compute gradients on ``step``, the
gradients are sent to a server runner which aggregates its inbox on its next ``step``.
Default aggregator = mean; Byzantine nodes are not modelled here."""
from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence

import torch

from ..node_cluster import NodeCluster


def _worker_fns(grad_fn: Callable[[], torch.Tensor]):
    def step(state: dict) -> dict:
        state["grad"] = grad_fn()
        return state

    def on_msg(state: dict, msg: Any) -> dict:
        return state

    return step, on_msg


def _server_fns(agg: Callable[[Sequence[torch.Tensor]], torch.Tensor]):
    def step(state: dict) -> dict:
        pending = state.get("in_msgs") or []
        if pending:
            state["out"] = agg(pending)
            state["in_msgs"] = []
        return state

    def on_msg(state: dict, msg: Any) -> dict:
        state.setdefault("in_msgs", []).append(msg)
        return state

    return step, on_msg


def _mean(grads: Sequence[torch.Tensor]) -> torch.Tensor:
    return sum(grads) / len(grads)


class ParameterServerRunner:
    def __init__(self, worker_grad_fns: List[Callable[[], torch.Tensor]],
                 aggregator: Optional[Callable[[Sequence[torch.Tensor]], torch.Tensor]] = None, *,
                 transport=None) -> None:
        self.cluster = NodeCluster(transport=transport)
        self.server_id = "server"
        self.worker_ids: List[str] = []
        self.cluster.add_node(self.server_id, *_server_fns(aggregator or _mean), init_state={})
        for idx, fn in enumerate(worker_grad_fns):
            wid = f"w{idx}"
            self.cluster.add_node(wid, *_worker_fns(fn), init_state={})
            self.worker_ids.append(wid)

    def start(self) -> None:
        self.cluster.start_all()

    def stop(self) -> None:
        self.cluster.stop_all()

    def run_round(self) -> torch.Tensor:
        for wid in self.worker_ids:
            self.cluster._nodes[wid].step()
        for wid in self.worker_ids:
            self.cluster.send(self.server_id, self.cluster.state(wid).get("grad"))
        self.cluster.barrier(0.01)  # lets TCP transports land their messages
        self.cluster._nodes[self.server_id].step()
        return self.cluster.state(self.server_id).get("out")


__all__ = ["ParameterServerRunner"]
