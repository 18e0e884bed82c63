"""Mean of medians (MeaMed): per coordinate, average the n-f values closest to the median
(reference aggregators/coordinate_wise/mean_of_medians.py:28-162)."""
from __future__ import annotations

from ... import ops
from ..base import CoordinateWiseAggregator


class MeanOfMedians(CoordinateWiseAggregator):
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
