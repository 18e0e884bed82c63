"""Gossip step functions mixed into user node classes (reference engine/node/mixin.py:27-105).

``P2PHonestMixin`` expects the host class to provide ``model``, ``device``, ``criterion``,
``next_batch()`` and a robust ``p2p_agg`` (optionally a ``p2p_pre`` pre-aggregator):

* ``p2p_half_step(lr)``: forward/backward on the next batch, plain SGD step in place, return the
  flat parameter vector theta^{t+1/2} that gets broadcast to the out-neighbours;
* ``p2p_aggregate_and_set(own, received)``: robustly aggregate own + received vectors and load the
  result into the model.

A host that binds its model to a :class:`~byzpy_b200.parallel.arena.ParamArena` (attribute ``arena``)
gets the flat fast path: parameters and gradients are views of two flat buffers, so the SGD half
step is one fused update, the broadcast vector one copy and loading the aggregate one copy,
instead of one small op per parameter tensor.

``P2PByzantineMixin.p2p_broadcast_vector`` is the adversary: it runs the node's ``attack`` on the
vectors it saw from honest neighbours.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from ...aggregators.base import Aggregator
from ...attacks.base import Attack
from ...parallel.arena import flatten_params, write_vector_to_params_
from ...pre_aggregators.base import PreAggregator

Tensor = torch.Tensor


class P2PHonestMixin:
    """The honest side of a gossip round, for nodes that keep a ``torch.nn.Module``.

    Mix it into a node class that provides ``model``, ``device``, ``criterion``, ``next_batch()``, a robust aggregator
    ``p2p_agg`` and optionally a pre-aggregator ``p2p_pre``.  A round is two calls:

    * :meth:`p2p_half_step` ``(lr)`` -- one local SGD step on the next batch; returns the flat parameter vector to
      broadcast;
    * :meth:`p2p_aggregate_and_set` ``(own, neighbour_vectors)`` -- robustly aggregate own and received vectors and
      load the result into the model.

    :meth:`get_param_vector` / :meth:`set_param_vector` expose the model as one flat vector.  When the node holds a
    :class:`~byzpy_b200.parallel.arena.ParamArena` over its model (attribute ``arena``), all four work on the arena's
    flat buffers without per-parameter copies.
    """

    model: nn.Module
    device: torch.device
    criterion: nn.Module
    optimizer: torch.optim.Optimizer
    p2p_agg: Aggregator
    p2p_pre: Optional[PreAggregator] = None

    def next_batch(self) -> Tuple[Tensor, Tensor]:  # supplied by the concrete node
        raise NotImplementedError

    # -- flat view of the model ---------------------------------------------------------------
    def _bound_arena(self):
        arena = getattr(self, "arena", None)
        return arena if arena is not None and arena.module is self.model and arena.check_bound() else None

    def get_param_vector(self) -> Tensor:
        arena = self._bound_arena()
        if arena is not None:
            return arena.param_vector().detach().clone()
        return flatten_params(self.model).to(self.device)

    def set_param_vector(self, vec: Tensor) -> None:
        arena = self._bound_arena()
        if arena is not None:
            with torch.no_grad():
                arena.param_vector().copy_(torch.as_tensor(vec).reshape(-1))
            return
        write_vector_to_params_(self.model, vec.to(self.device))

    # -- the two halves of a gossip round -----------------------------------------------------
    def p2p_half_step(self, lr: float) -> Tensor:
        inputs, targets = self.next_batch()
        arena = self._bound_arena()
        if arena is not None:
            arena.zero_grad()
            self.criterion(self.model(inputs), targets).backward()
            with torch.no_grad():
                arena.flat_params.sub_(arena.flat_grads, alpha=lr)
            return arena.param_vector().detach().clone()
        self.model.zero_grad(set_to_none=True)
        loss = self.criterion(self.model(inputs), targets)
        loss.backward()
        with torch.no_grad():
            for p in self.model.parameters():
                if p.grad is not None:
                    p.sub_(p.grad, alpha=lr)
        return self.get_param_vector()

    def p2p_aggregate_and_set(self, self_theta_half: Tensor, neighbor_vectors: List[Tensor]) -> None:
        candidates = [self_theta_half, *neighbor_vectors]
        if self.p2p_pre is not None:
            candidates = self.p2p_pre.pre_aggregate(candidates)
        self.set_param_vector(self.p2p_agg.aggregate(candidates))


class P2PByzantineMixin:
    """The Byzantine side of a gossip round: :meth:`p2p_broadcast_vector` hands the vectors received from honest
    neighbours to ``self.attack`` (as ``honest_grads``) and returns the crafted vector, cast to the dtype and device of
    ``like``.
    """

    device: torch.device
    attack: Attack

    def p2p_broadcast_vector(self, *, neighbor_vectors: Optional[List[Tensor]] = None,
                             like: Optional[Tensor] = None) -> Tensor:
        crafted = self.attack.apply(model=None, x=None, y=None, honest_grads=neighbor_vectors, base_grad=None)
        crafted = torch.as_tensor(crafted)
        return crafted if like is None else crafted.to(device=like.device, dtype=like.dtype)


__all__ = ["P2PHonestMixin", "P2PByzantineMixin"]
