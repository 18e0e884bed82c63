"""Inf attack: an all ``+inf`` vector (reference attacks/inf.py:35-119)."""
from __future__ import annotations

import torch

from .. import ops
from ..aggregators._chunking import select_adaptive_chunk_size
from ..aggregators.base import finish, pool_size_of, prepare_rows
from ..engine.graph.subtask import SubTask
from .base import Attack


def _inf_chunk(start: int, end: int, device: str):
    out = torch.empty(end - start, dtype=torch.float32, device=device)
    ops.fill_(out, float("inf"))
    return start, out


class InfAttack(Attack):
    """Inf attack: submit a vector of ``+inf`` shaped like a gradient.

    Breaks any aggregator that averages without screening (the mean becomes ``inf``); selection-based aggregators
    order ``+inf`` as the largest value and discard it.

    Parameters
    ----------
    chunk_size : int, default 8192
        Coordinates per subtask on an actor pool.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.attacks import InfAttack
    >>> InfAttack().apply(honest_grads=[torch.zeros(3)])
    tensor([inf, inf, inf])
    """

    name = "inf"
    max_subtasks_inflight = 0       # 0 / None: the pool-sized default window (value of the reference class)
    uses_honest_grads = True
    supports_subtasks = True

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

    # -- subtask path: every chunk of the output is filled independently (reference inf.py:29-32, 80-117) --
    def create_subtasks(self, inputs, *, context):
        grads = inputs.get("honest_grads")
        if not grads:
            raise ValueError("InfAttack requires honest_grads.")
        rows, like = prepare_rows([grads[0]], "honest_grads")
        d = rows[0].numel()
        chunk = select_adaptive_chunk_size(d, self.chunk_size, pool_size=pool_size_of(context))
        return [SubTask(fn=_inf_chunk, args=(s, min(d, s + chunk), str(like.device)), name=f"inf_chunk_{k}")
                for k, s in enumerate(range(0, d, chunk))]

    def reduce_subtasks(self, partials, inputs, *, context):
        if not partials:
            return self.compute(inputs, context=context)
        _, like = prepare_rows([inputs["honest_grads"][0]], "honest_grads")
        parts = sorted(partials, key=lambda p: p[0])
        return finish(torch.cat([torch.as_tensor(p[1]).reshape(-1).to(like.device) for p in parts]), like)


__all__ = ["InfAttack"]
