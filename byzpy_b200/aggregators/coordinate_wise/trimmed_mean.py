"""Coordinate-wise trimmed mean: per coordinate drop the f smallest and f largest of the n
values and average the rest (reference aggregators/coordinate_wise/trimmed_mean.py:27-211)."""
from __future__ import annotations

from ... import ops
from ..base import CoordinateWiseAggregator


class CoordinateWiseTrimmedMean(CoordinateWiseAggregator):
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
