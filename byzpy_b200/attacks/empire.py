"""Fall-of-empires / inner-product manipulation: ``scale * mean(honest_grads)``
(reference attacks/empire.py:23-187)."""
from __future__ import annotations

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


__all__ = ["EmpireAttack"]
