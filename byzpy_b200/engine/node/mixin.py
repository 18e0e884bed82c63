"""Peer-to-peer step functions mixed into user node classes
(reference engine/node/mixin.py:27-105): ``p2p_half_step`` (fwd/bwd + in-place SGD, returns the
flat parameters), ``p2p_aggregate_and_set`` (optional pre-aggregation, robust aggregation of own +
neighbour vectors, write back), and the Byzantine ``p2p_broadcast_vector``."""
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
    model: nn.Module
    device: torch.device
    criterion: nn.Module
    optimizer: torch.optim.Optimizer
    p2p_agg: Aggregator
    p2p_pre: Optional[PreAggregator] = None

    def next_batch(self) -> Tuple[Tensor, Tensor]:  # provided by the concrete node
        raise NotImplementedError

    def get_param_vector(self) -> Tensor:
        return flatten_params(self.model).to(self.device)

    def set_param_vector(self, vec: Tensor) -> None:
        write_vector_to_params_(self.model, vec.to(self.device))

    def p2p_half_step(self, lr: float) -> Tensor:
        x, y = self.next_batch()
        self.model.zero_grad(set_to_none=True)
        self.criterion(self.model(x), y).backward()
        with torch.no_grad():
            for p in self.model.parameters():
                if p.grad is not None:
                    p.add_(p.grad, alpha=-lr)
        return self.get_param_vector()

    def p2p_aggregate_and_set(self, self_theta_half: Tensor, neighbor_vectors: List[Tensor]) -> None:
        vecs = [self_theta_half] + list(neighbor_vectors)
        if self.p2p_pre is not None:
            vecs = self.p2p_pre.pre_aggregate(vecs)
        self.set_param_vector(self.p2p_agg.aggregate(vecs))


class P2PByzantineMixin:
    device: torch.device
    attack: Attack

    def p2p_broadcast_vector(self, *, neighbor_vectors: Optional[List[Tensor]] = None,
                             like: Optional[Tensor] = None) -> Tensor:
        out = torch.as_tensor(self.attack.apply(model=None, x=None, y=None,
                                                honest_grads=neighbor_vectors, base_grad=None))
        if like is not None:
            out = out.to(device=like.device, dtype=like.dtype)
        return out


__all__ = ["P2PHonestMixin", "P2PByzantineMixin"]
