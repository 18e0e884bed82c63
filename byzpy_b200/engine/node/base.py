"""The three node contracts the orchestrators program against (reference engine/node/base.py:9-39).

* ``Node``: owns a data stream (``next_batch``) and accepts the server's aggregated gradient;
* ``HonestNode``: turns a batch into a gradient;
* ``ByzantineNode``: produces whatever it wants, optionally after seeing the honest gradients
  (omniscient adversary).  Byzantine nodes own no data by default, so the ``*_for_next_batch``
  helper feeds them an empty batch.

The ``*_for_next_batch`` helpers are what ``ParameterServer.round()`` calls through the actor
layer; they return detached tensors so nothing drags an autograd graph across an actor boundary.
"""
from __future__ import annotations

import abc
from typing import List, Optional, Tuple

import torch

Batch = Tuple[torch.Tensor, torch.Tensor]


class Node(abc.ABC):
    @abc.abstractmethod
    def next_batch(self) -> Batch:
        """The node's next (inputs, targets) mini-batch."""

    @abc.abstractmethod
    def apply_server_gradient(self, grad_vec: torch.Tensor) -> None:
        """Apply the flat aggregated gradient the parameter server broadcast."""


class HonestNode(Node):
    @abc.abstractmethod
    def honest_gradient(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """Flat gradient of the local loss on ``(x, y)``."""

    def honest_gradient_for_next_batch(self) -> torch.Tensor:
        batch = self.next_batch()
        return self.honest_gradient(*batch).detach()


class ByzantineNode(Node):
    @abc.abstractmethod
    def byzantine_gradient(self, x: torch.Tensor, y: torch.Tensor,
                           honest_grads: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        """The vector this adversary submits instead of a gradient."""

    def byzantine_gradient_for_next_batch(self, honest_grads: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        no_x, no_y = torch.empty(0), torch.empty(0, dtype=torch.long)
        return self.byzantine_gradient(no_x, no_y, honest_grads=honest_grads).detach()


__all__ = ["Node", "HonestNode", "ByzantineNode"]
