"""Sign-flip attack: ``scale * base_grad`` (reference attacks/sign_flip.py:22-145).  In the fused
device round this is a per-row scale folded into the aggregation kernel's load."""
from __future__ import annotations


import torch

from .. import ops
from ..aggregators._chunking import select_adaptive_chunk_size
from ..aggregators.base import finish, pool_size_of, prepare_rows
from ..engine.graph.subtask import SubTask
from .base import Attack


def _scale_chunk(vec: torch.Tensor, start: int, end: int, scale: float):
    return start, ops.scale_copy(vec[start:end], scale)


class SignFlipAttack(Attack):
    """Sign flip: submit the node's own gradient multiplied by ``scale`` (default ``-1``).

    Parameters
    ----------
    scale : float, default -1.0
        Multiplier of the node's gradient.
    chunk_size : int, default 8192
        Coordinates per subtask on an actor pool.

    Notes
    -----
    Needs ``base_grad`` only (not omniscient).  In the fused device round it is a per-row scale applied as the
    aggregation kernel loads the row; this is the attack of the headline benchmark (6 honest + 2 sign-flipping
    workers).

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.attacks import SignFlipAttack
    >>> SignFlipAttack().apply(base_grad=torch.tensor([1.0, -2.0]))
    tensor([-1.,  2.])
    """

    name = "sign-flip"
    max_subtasks_inflight = 0       # 0 / None: the pool-sized default window (value of the reference class)
    uses_base_grad = True
    supports_subtasks = True

    def __init__(self, scale: float = -1.0, *, chunk_size: int = 8192) -> None:
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.scale = float(scale)
        self.chunk_size = int(chunk_size)

    def apply(self, *, model=None, x=None, y=None, honest_grads=None, base_grad=None):
        if base_grad is None:
            raise ValueError("SignFlipAttack requires base_grad.")
        rows, like = prepare_rows([base_grad], "base_grad")
        return finish(ops.scale_copy(rows[0], self.scale), like)

    def fold(self, n_honest: int):
        from ..parallel.device_ps import RowFold

        return RowFold("scale", scale=self.scale)

    def create_subtasks(self, inputs, *, context):
        base = inputs.get("base_grad")
        if base is None:
            return []
        rows, _ = prepare_rows([base], "base_grad")
        vec = rows[0]
        d = vec.numel()
        chunk = select_adaptive_chunk_size(d, self.chunk_size, pool_size=pool_size_of(context))
        return [SubTask(fn=_scale_chunk, args=(vec, s, min(d, s + chunk), self.scale),
                        name=f"signflip_chunk_{k}") for k, s in enumerate(range(0, d, chunk))]

    def reduce_subtasks(self, partials, inputs, *, context):
        if not partials:
            return self.compute(inputs, context=context)
        _, like = prepare_rows([inputs["base_grad"]], "base_grad")
        parts = sorted(partials, key=lambda p: p[0])
        return finish(torch.cat([p[1].reshape(-1).to(like.device) for p in parts]), like)


__all__ = ["SignFlipAttack"]
