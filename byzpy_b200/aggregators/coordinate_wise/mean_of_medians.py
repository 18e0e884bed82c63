"""Mean of medians (MeaMed): per coordinate, average the n-f values closest to the median
(reference aggregators/coordinate_wise/mean_of_medians.py:28-162)."""
from __future__ import annotations

from ... import ops
from ..base import CoordinateWiseAggregator


class MeanOfMedians(CoordinateWiseAggregator):
    """MeaMed: per coordinate, the mean of the ``n - f`` values closest to that coordinate's median.

    Parameters
    ----------
    f : int
        Number of values discarded per coordinate (the ones farthest from the median); ``0 <= f < n``.
    chunk_size : int, default 8192
        Coordinates per subtask on an actor pool.

    Notes
    -----
    Ties in the distance to the median are broken by sorted position, which is what makes the result
    deterministic on every device.  CUDA inputs: the sorted window of ``n - f`` consecutive order statistics with the
    smallest spread around the median is found inside the same selection-network launch as the median.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.coordinate_wise import MeanOfMedians
    >>> grads = [torch.tensor([v]) for v in (1.0, 2.0, 3.0, 50.0)]
    >>> MeanOfMedians(f=1).aggregate(grads)
    tensor([2.])
    """

    name = "mean-of-medians"
    _mode = ops.MODE_MEAMED

    def __init__(self, f: int, *, chunk_size: int = 8192) -> None:
        if f < 0:
            raise ValueError("f must be >= 0")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.f = int(f)
        self.chunk_size = int(chunk_size)

    def _f(self, n: int) -> int:
        return self.f

    def _validate(self, n: int) -> None:
        if not (0 <= self.f < n):
            raise ValueError(f"f must satisfy 0 <= f < n (got n={n}, f={self.f})")


__all__ = ["MeanOfMedians"]
