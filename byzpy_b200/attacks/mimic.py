"""Mimic attack: replay honest gradient number ``epsilon`` unchanged (reference
attacks/mimic.py:35-142).  In the fused device round this is a row *alias* (same pointer twice)."""
from __future__ import annotations

from .. import ops
from ..aggregators.base import finish, prepare_rows
from .base import Attack


class MimicAttack(Attack):
    name = "mimic"
    uses_honest_grads = True
    supports_subtasks = False

    def __init__(self, epsilon: int = 0, *, chunk_size: int = 8192) -> None:
        if epsilon < 0:
            raise ValueError("epsilon must be >= 0")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.epsilon = int(epsilon)
        self.chunk_size = int(chunk_size)

    def apply(self, *, model=None, x=None, y=None, honest_grads=None, base_grad=None):
        if not honest_grads:
            raise ValueError("MimicAttack requires honest_grads.")
        if self.epsilon >= len(honest_grads):
            raise ValueError(f"epsilon={self.epsilon} out of range for {len(honest_grads)} honest gradients")
        rows, like = prepare_rows([honest_grads[self.epsilon]], "honest_grads")
        return finish(ops.scale_copy(rows[0], 1.0), like)

    def fold(self, n_honest: int):
        from ..parallel.device_ps import RowFold

        return RowFold("alias", index=self.epsilon)


__all__ = ["MimicAttack"]
