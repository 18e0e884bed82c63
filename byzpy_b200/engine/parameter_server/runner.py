"""``ParameterServerRunner``: the reference's process-per-node parameter-server prototype
(reference engine/parameter_server/runner.py:49-90) on top of :class:`NodeCluster`.

One ``NodeRunner`` process per worker plus one for the server.  A round is driven from the parent:

1. every worker executes one ``step`` -> its state holds a fresh gradient;
2. the parent forwards each gradient to the server's inbox (directly or through the transport);
3. the server executes one ``step`` -> it folds its inbox with ``aggregator`` (mean by default).

Byzantine behaviour is not modelled at this level (same as the reference); use
:class:`~byzpy_b200.engine.parameter_server.ps.ParameterServer` for real training.
"""
from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence

import torch

from ..node_cluster import NodeCluster

GradFn = Callable[[], torch.Tensor]
AggFn = Callable[[Sequence[torch.Tensor]], torch.Tensor]

SERVER = "server"


def mean_aggregate(grads: Sequence[torch.Tensor]) -> torch.Tensor:
    """Arithmetic mean of the gradients: the default (non-robust) aggregator of the prototype runner."""
    total = grads[0].clone()
    for g in grads[1:]:
        total += g
    return total / len(grads)


class _Worker:
    """Picklable (step, on_message) pair of a worker runner."""

    def __init__(self, grad_fn: GradFn):
        self.grad_fn = grad_fn

    def step(self, state: dict) -> dict:
        state["grad"] = self.grad_fn()
        return state

    @staticmethod
    def on_message(state: dict, msg: Any) -> dict:
        return state                      # workers ignore their inbox in this prototype


class _Server:
    def __init__(self, aggregate: AggFn):
        self.aggregate = aggregate

    def step(self, state: dict) -> dict:
        pending = state.get("in_msgs")
        if pending:
            state["out"] = self.aggregate(pending)
            state["in_msgs"] = []
        return state

    @staticmethod
    def on_message(state: dict, msg: Any) -> dict:
        state.setdefault("in_msgs", []).append(msg)
        return state


class ParameterServerRunner:
    """A prototype parameter server built from blocking :class:`~byzpy_b200.engine.node_runner.NodeRunner` processes.

    One server process and one process per worker.  ``run_round()`` steps every worker (each calls its gradient
    function), forwards the gradients to the server, steps the server (which aggregates its inbox) and returns the
    result.

    Parameters
    ----------
    worker_grad_fns : list of callables
        One ``() -> tensor`` per worker (shipped to the worker process by value).
    aggregator : callable, optional
        ``sequence of tensors -> tensor``; default :func:`mean_aggregate`.
    transport : Transport, optional
        ``LocalTransport()`` / ``TcpTransport()`` to route messages through a transport instead of the runners' pipes.

    Notes
    -----
    ``start()`` / ``stop()`` bracket the rounds.  ``examples/ps/decentralized_demo.py`` is a complete program.
    """

    def __init__(self, worker_grad_fns: List[GradFn], aggregator: Optional[AggFn] = None, *,
                 transport=None) -> None:
        self.cluster = NodeCluster(transport=transport)
        self.server_id = SERVER
        server = _Server(aggregator or mean_aggregate)
        self.cluster.add_node(SERVER, server.step, server.on_message, init_state={})
        self.worker_ids: List[str] = []
        for k, fn in enumerate(worker_grad_fns):
            worker = _Worker(fn)
            self.cluster.add_node(f"w{k}", worker.step, worker.on_message, init_state={})
            self.worker_ids.append(f"w{k}")

    def start(self) -> None:
        self.cluster.start_all()

    def stop(self) -> None:
        self.cluster.stop_all()

    def run_round(self) -> torch.Tensor:
        nodes = self.cluster._nodes
        for wid in self.worker_ids:
            nodes[wid].step()
        for wid in self.worker_ids:
            self.cluster.send(SERVER, self.cluster.state(wid).get("grad"))
        self.cluster.barrier(0.01)            # lets asynchronous (TCP) transports land their frames
        nodes[SERVER].step()
        return self.cluster.state(SERVER).get("out")


__all__ = ["ParameterServerRunner", "mean_aggregate"]
