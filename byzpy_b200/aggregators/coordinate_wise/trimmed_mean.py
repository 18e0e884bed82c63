"""Coordinate-wise trimmed mean: per coordinate drop the f smallest and f largest of the n
values and average the rest (reference aggregators/coordinate_wise/trimmed_mean.py:27-211)."""
from __future__ import annotations

from ... import ops
from ..base import CoordinateWiseAggregator


class CoordinateWiseTrimmedMean(CoordinateWiseAggregator):
    """Mean of every coordinate after dropping its ``f`` smallest and ``f`` largest values.

    Parameters
    ----------
    f : int
        Values trimmed from each end, per coordinate; needs ``0 <= 2 f < n``.
    chunk_size : int, default 8192
        Coordinates per subtask on an actor pool.

    Notes
    -----
    CUDA inputs: full odd-even merge network per coordinate followed by a masked sum, one launch
    (``cw_select_staged_kernel``: each thread streams its next tiles through a private cp.async ring).  ``f = 0`` is the
    plain mean.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseTrimmedMean
    >>> grads = [torch.tensor([v]) for v in (0.0, 1.0, 2.0, 3.0, 1000.0)]
    >>> CoordinateWiseTrimmedMean(f=1).aggregate(grads)
    tensor([2.])
    """

    name = "coordinate-wise-trimmed-mean"
    _mode = ops.MODE_TRMEAN

    def __init__(self, f: int, *, chunk_size: int = 4096) -> None:
        if f < 0:
            raise ValueError("f must be >= 0")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.f = int(f)
        self.chunk_size = int(chunk_size)

    def _f(self, n: int) -> int:
        return self.f

    def _validate(self, n: int) -> None:
        if not (0 <= 2 * self.f < n):
            raise ValueError(f"f must satisfy 0 <= 2f < n (got n={n}, f={self.f})")


__all__ = ["CoordinateWiseTrimmedMean"]
