"""Nearest-neighbour mixing: replace every vector by the mean of its ``n - f`` nearest
vectors, itself included (reference pre_aggregators/nnm.py:21-194)."""
from __future__ import annotations

from ..ops import nspace
from .base import LinearPreAggregator


class NearestNeighborMixing(LinearPreAggregator):
    name = "pre-agg/nnm"

    def __init__(self, f: int, *, feature_chunk_size: int = 8192) -> None:
        if f < 0:
            raise ValueError("f must be >= 0")
        if feature_chunk_size <= 0:
            raise ValueError("feature_chunk_size must be > 0")
        self.f = int(f)
        self.feature_chunk_size = int(feature_chunk_size)

    def _validate(self, n: int) -> None:
        if not (0 <= self.f < n):
            raise ValueError(f"f must satisfy 0 <= f < n (got n={n}, f={self.f})")

    def row_map(self, G, n):
        return nspace.nnm_matrix(G, self.f)


    def row_map_device(self, G, n):
        from ..ops import nspace_cuda

        return nspace_cuda.nnm_matrix(G, self.f)

__all__ = ["NearestNeighborMixing"]
