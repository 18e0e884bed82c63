"""Node contracts used by the parameter-server and peer-to-peer orchestrators
(reference engine/node/base.py:9-39)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Optional, Tuple

import torch


class Node(ABC):
    @abstractmethod
    def next_batch(self) -> Tuple[torch.Tensor, torch.Tensor]:
        ...

    @abstractmethod
    def apply_server_gradient(self, grad_vec: torch.Tensor) -> None:
        ...


class HonestNode(Node, ABC):
    @abstractmethod
    def honest_gradient(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        ...

    def honest_gradient_for_next_batch(self) -> torch.Tensor:
        x, y = self.next_batch()
        return self.honest_gradient(x, y).detach()


class ByzantineNode(Node, ABC):
    @abstractmethod
    def byzantine_gradient(self, x: torch.Tensor, y: torch.Tensor,
                           honest_grads: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        ...

    def byzantine_gradient_for_next_batch(self, honest_grads: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        # Byzantine nodes own no data by default: empty batch, as in the reference contract.
        return self.byzantine_gradient(torch.empty(0), torch.empty(0, dtype=torch.long),
                                       honest_grads=honest_grads).detach()


__all__ = ["Node", "HonestNode", "ByzantineNode"]
