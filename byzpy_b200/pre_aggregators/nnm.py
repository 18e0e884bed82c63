"""Nearest-neighbour mixing: replace every vector by the mean of its ``n - f`` nearest
vectors, itself included (reference pre_aggregators/nnm.py:21-194)."""
from __future__ import annotations

from ..ops import nspace
from .base import LinearPreAggregator


class NearestNeighborMixing(LinearPreAggregator):
    """Nearest-Neighbour Mixing: replace every vector by the mean of its ``n - f`` nearest vectors (itself included).

    Honest vectors end up averaged with mostly honest neighbours, which shrinks their spread; whatever a Byzantine
    vector is replaced by is an average dominated by honest ones.

    Parameters
    ----------
    f : int
        Expected number of Byzantine vectors; ``0 <= f < n``.
    feature_chunk_size : int, default 8192
        Coordinates per subtask on an actor pool.

    Notes
    -----
    ``X' = W X`` with a 0/1 neighbour matrix ``W / (n - f)`` read off the distance matrix: one Gram pass, the
    ``n x n`` k-nearest selection on one CTA (ties to the lower index), and one pass of the multi-row weighted-sum
    kernel (``csrc/wsum.cu``) that reads every input once for all ``n`` outputs.  When a Gram-family aggregator
    follows, the mixed vectors are never written: its Gram matrix is ``W G W^T``.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.pre_aggregators import NearestNeighborMixing
    >>> xs = [torch.tensor([0.0]), torch.tensor([1.0]), torch.tensor([2.0]), torch.tensor([90.0])]
    >>> NearestNeighborMixing(f=1).pre_aggregate(xs)
    [tensor([1.]), tensor([1.]), tensor([1.]), tensor([31.])]
    """

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
