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
    """What every participant of a training run provides: a stream of batches and a way to take the agreed update.

    Implement :meth:`next_batch` (the node's next ``(inputs, targets)`` mini-batch) and
    :meth:`apply_server_gradient` (apply the flat aggregated gradient: write it into ``.grad`` and step an optimizer, or
    update a flat parameter buffer directly).  Orchestrators only call these through an actor proxy, so a node may live in
    a thread, another process, another machine or on a CUDA stream.
    """

    @abc.abstractmethod
    def next_batch(self) -> Batch:
        """The node's next (inputs, targets) mini-batch."""

    @abc.abstractmethod
    def apply_server_gradient(self, grad_vec: torch.Tensor) -> None:
        """Apply the flat aggregated gradient the parameter server broadcast."""


class HonestNode(Node):
    """A node that follows the protocol: it turns a batch into the gradient of its local loss.

    Implement :meth:`honest_gradient` ``(x, y) -> flat gradient`` in ``model.parameters()`` order.
    ``honest_gradient_for_next_batch`` -- what :class:`~byzpy_b200.engine.parameter_server.ps.ParameterServer`
    calls each round -- draws the batch itself and returns the gradient detached.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.engine.node.base import HonestNode
    >>> class Quadratic(HonestNode):
    ...     # minimises |w - target|^2 / 2; its gradient is w - target
    ...     def __init__(self, target):
    ...         self.w, self.target = torch.zeros(2), torch.as_tensor(target)
    ...     def next_batch(self):
    ...         return self.target, torch.empty(0)
    ...     def honest_gradient(self, x, y):
    ...         return self.w - x
    ...     def apply_server_gradient(self, g):
    ...         self.w -= 0.5 * g
    >>> node = Quadratic([2.0, 4.0])
    >>> node.apply_server_gradient(node.honest_gradient_for_next_batch()); node.w
    tensor([1., 2.])
    """

    @abc.abstractmethod
    def honest_gradient(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """Flat gradient of the local loss on ``(x, y)``."""

    def honest_gradient_for_next_batch(self) -> torch.Tensor:
        batch = self.next_batch()
        return self.honest_gradient(*batch).detach()


class ByzantineNode(Node):
    """A node controlled by the adversary: it submits whatever :meth:`byzantine_gradient` returns.

    ``byzantine_gradient(x, y, honest_grads=None)`` receives the honest gradients of the round when the orchestrator
    runs the omniscient setting (the parameter server does), and usually delegates to an
    :class:`~byzpy_b200.attacks.base.Attack`.  Byzantine nodes own no data by default:
    ``byzantine_gradient_for_next_batch`` passes an empty batch.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.attacks import EmpireAttack
    >>> from byzpy_b200.engine.node.base import ByzantineNode
    >>> class Empire(ByzantineNode):
    ...     attack = EmpireAttack(scale=-1.0)
    ...     def next_batch(self):
    ...         return torch.empty(0), torch.empty(0, dtype=torch.long)
    ...     def apply_server_gradient(self, g):
    ...         pass
    ...     def byzantine_gradient(self, x, y, honest_grads=None):
    ...         return self.attack.apply(honest_grads=honest_grads)
    >>> Empire().byzantine_gradient_for_next_batch([torch.tensor([1.0]), torch.tensor([3.0])])
    tensor([-2.])
    """

    @abc.abstractmethod
    def byzantine_gradient(self, x: torch.Tensor, y: torch.Tensor,
                           honest_grads: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        """The vector this adversary submits instead of a gradient."""

    def byzantine_gradient_for_next_batch(self, honest_grads: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        no_x, no_y = torch.empty(0), torch.empty(0, dtype=torch.long)
        return self.byzantine_gradient(no_x, no_y, honest_grads=honest_grads).detach()


__all__ = ["Node", "HonestNode", "ByzantineNode"]
