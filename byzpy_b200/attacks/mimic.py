"""Mimic attack: replay honest gradient number ``epsilon`` unchanged (reference
attacks/mimic.py:35-142).  In the fused device round this is a row *alias* (same pointer twice)."""
from __future__ import annotations

import torch

from .. import ops
from ..aggregators._chunking import select_adaptive_chunk_size
from ..aggregators.base import finish, pool_size_of, prepare_rows
from ..engine.graph.subtask import SubTask
from .base import Attack


def _copy_chunk(vec: torch.Tensor, start: int, end: int):
    return start, ops.scale_copy(vec[start:end], 1.0)


class MimicAttack(Attack):
    """Mimic: replay the gradient of honest node number ``epsilon`` unchanged.

    Over-representing one honest node biases the aggregate towards that node's data when the honest data is
    heterogeneous, without ever submitting an outlier.

    Parameters
    ----------
    epsilon : int, default 0
        Index of the honest gradient to copy.
    chunk_size : int, default 8192
        Coordinates per subtask on an actor pool.

    Notes
    -----
    In the fused device round the Byzantine row is an alias of the victim's row (the same pointer twice): no copy.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.attacks import MimicAttack
    >>> MimicAttack(epsilon=1).apply(honest_grads=[torch.tensor([1.0]), torch.tensor([5.0])])
    tensor([5.])
    """

    name = "mimic"
    max_subtasks_inflight = 0       # 0 / None: the pool-sized default window (value of the reference class)
    uses_honest_grads = True
    supports_subtasks = True

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

    # -- subtask path: the replayed row is copied chunk by chunk (reference mimic.py:29-32, 100-140) -----
    def _victim(self, inputs):
        grads = inputs.get("honest_grads")
        if not grads:
            raise ValueError("MimicAttack requires honest_grads.")
        if self.epsilon >= len(grads):
            raise ValueError(f"epsilon={self.epsilon} out of range for {len(grads)} honest gradients")
        return prepare_rows([grads[self.epsilon]], "honest_grads")

    def create_subtasks(self, inputs, *, context):
        rows, _ = self._victim(inputs)
        vec = rows[0]
        d = vec.numel()
        chunk = select_adaptive_chunk_size(d, self.chunk_size, pool_size=pool_size_of(context))
        return [SubTask(fn=_copy_chunk, args=(vec, s, min(d, s + chunk)), name=f"mimic_chunk_{k}")
                for k, s in enumerate(range(0, d, chunk))]

    def reduce_subtasks(self, partials, inputs, *, context):
        if not partials:
            return self.compute(inputs, context=context)
        _, like = self._victim(inputs)
        parts = sorted(partials, key=lambda p: p[0])
        return finish(torch.cat([torch.as_tensor(p[1]).reshape(-1).to(like.device) for p in parts]), like)

    def fold(self, n_honest: int):
        from ..parallel.device_ps import RowFold

        return RowFold("alias", index=self.epsilon)


__all__ = ["MimicAttack"]
