"""``DecentralizedParameterServer``: the async facade the reference keeps for old scripts
(reference engine/parameter_server/decentralized.py:14-38).  It adapts node objects exposing a
``.grad`` attribute to the process-per-node :class:`ParameterServerRunner`; Byzantine nodes are
accepted for signature compatibility and, as in the reference, not simulated by this prototype.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from .runner import ParameterServerRunner

Aggregate = Callable[[Sequence[torch.Tensor]], torch.Tensor]


def _grad_reader(node) -> Callable[[], torch.Tensor]:
    def read() -> torch.Tensor:
        return node.grad
    return read


class DecentralizedParameterServer:
    """Async facade over the process-per-node :class:`~byzpy_b200.engine.parameter_server.runner.ParameterServerRunner`
    (kept for scripts written against the reference's prototype).

    Parameters
    ----------
    honest_nodes : list
        Objects exposing the gradient to submit as ``.grad``.
    byzantine_nodes : list, optional
        Accepted for signature compatibility; the prototype does not simulate them.
    aggregator : callable
        ``sequence of tensors -> tensor``.

    Notes
    -----
    ``await bootstrap()`` starts the processes, ``await round()`` returns the aggregate of one round,
    ``await shutdown()`` stops them.  For real training use
    :class:`~byzpy_b200.engine.parameter_server.ps.ParameterServer`.
    """

    def __init__(self, honest_nodes: List, byzantine_nodes: Optional[List], aggregator: Aggregate) -> None:
        self._honest = list(honest_nodes)
        self._byz = list(byzantine_nodes or [])
        self._runner = ParameterServerRunner(worker_grad_fns=[_grad_reader(h) for h in self._honest],
                                             aggregator=aggregator)
        self.rounds = 0

    @property
    def runner(self) -> ParameterServerRunner:
        return self._runner

    async def bootstrap(self) -> None:
        self._runner.start()

    async def round(self) -> torch.Tensor:
        out = self._runner.run_round()
        self.rounds += 1
        return out

    async def shutdown(self) -> None:
        self._runner.stop()


__all__ = ["DecentralizedParameterServer"]
