"""Fall-of-empires / inner-product manipulation: ``scale * mean(honest_grads)``
(reference attacks/empire.py:23-187)."""
from __future__ import annotations

from ..aggregators._chunking import select_adaptive_chunk_size
from ..aggregators.base import pool_size_of
from .base import ColumnStatAttack


class EmpireAttack(ColumnStatAttack):
    """Fall of Empires (inner-product manipulation): submit ``scale`` times the mean of the honest gradients.

    With a negative ``scale`` the submitted vector points against the honest descent direction; small magnitudes
    (``-0.1``) slip past distance-based defences while still flipping the sign of the aggregate's inner product with
    the true gradient.

    Parameters
    ----------
    scale : float, default -1.0
        Multiplier of the honest mean.
    chunk_size : int, default 8
        Honest gradients per subtask on an actor pool (partial sums).

    Notes
    -----
    Needs ``honest_grads``.  In the fused device round the row is synthesised per coordinate from the column mean the
    aggregation kernel computes anyway (``RowFold("virtual", a=scale, b=0)``: row = ``a * mean + b * std``).

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.attacks import EmpireAttack
    >>> EmpireAttack(scale=-2.0).apply(honest_grads=[torch.tensor([1.0, 2.0]), torch.tensor([3.0, 4.0])])
    tensor([-4., -6.])
    """

    name = "empire"

    def __init__(self, scale: float = -1.0, *, chunk_size: int = 8) -> None:
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.scale = float(scale)
        self.chunk_size = int(chunk_size)

    def _coeffs(self, n_honest: int):
        return self.scale, 0.0

    def _subtask_feature_chunk(self, d: int, n_rows: int, context) -> int:
        # ``chunk_size`` counts GRADIENTS per subtask (reference attacks/empire.py:108-120: partial sums over row
        # blocks).  The column-statistics kernel splits coordinates instead, so the setting is honoured as a number
        # of subtasks: as many as the reference would create, but at least one per worker once there is enough
        # data for that.  (Read as a coordinate count, the default of 8 would mean d / 8 subtasks.)
        pool = pool_size_of(context)
        rows_chunk = max(1, select_adaptive_chunk_size(n_rows, self.chunk_size, pool_size=pool))
        pieces = -(-n_rows // rows_chunk)
        if d >= 4096 * max(1, pool):
            pieces = max(pieces, pool)
        per = -(-d // max(1, pieces))
        # cache-line multiples (16 floats) once there is a line per piece; below that the reference's subtask COUNT
        # is what callers observe (its test suite compares the counts for two pool sizes on a 256 x 256 input)
        return -(-per // 16) * 16 if per >= 64 else max(4, -(-per // 4) * 4)


__all__ = ["EmpireAttack"]
