"""Inf attack: an all ``+inf`` vector (reference attacks/inf.py:35-119)."""
from __future__ import annotations

import torch

from .. import ops
from ..aggregators.base import finish, prepare_rows
from .base import Attack


class InfAttack(Attack):
    name = "inf"
    uses_honest_grads = True
    supports_subtasks = False

    def __init__(self, *, chunk_size: int = 8192) -> None:
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.chunk_size = int(chunk_size)

    def apply(self, *, model=None, x=None, y=None, honest_grads=None, base_grad=None):
        if not honest_grads:
            raise ValueError("InfAttack requires honest_grads.")
        rows, like = prepare_rows([honest_grads[0]], "honest_grads")
        out = torch.empty(rows[0].numel(), dtype=torch.float32, device=like.device)
        ops.fill_(out, float("inf"))
        return finish(out, like)


__all__ = ["InfAttack"]
