"""Backwards-compatible wrapper around :class:`ParameterServerRunner`
(reference engine/parameter_server/decentralized.py:14-38)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from .runner import ParameterServerRunner


class DecentralizedParameterServer:
    def __init__(self, honest_nodes: List, byzantine_nodes: Optional[List],
                 aggregator: Callable[[Sequence[torch.Tensor]], torch.Tensor]) -> None:
        self._honest = honest_nodes
        self._byz = byzantine_nodes or []
        self._runner = ParameterServerRunner(
            worker_grad_fns=[(lambda h=h: h.grad) for h in self._honest], aggregator=aggregator)

    async def bootstrap(self) -> None:
        self._runner.start()

    async def round(self) -> torch.Tensor:
        return self._runner.run_round()

    async def shutdown(self) -> None:
        self._runner.stop()


__all__ = ["DecentralizedParameterServer"]
