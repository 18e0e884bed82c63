"""Bucketing: permute the inputs, split into consecutive buckets of ``bucket_size`` and return
the bucket means (reference pre_aggregators/bucketing.py:28-248).  Without ``perm`` the order
comes from ``rng`` (a fresh OS-seeded ``random.Random`` by default, like the reference)."""
from __future__ import annotations

import random
from typing import Iterable, List, Optional

from ..ops import nspace
from .base import LinearPreAggregator


class Bucketing(LinearPreAggregator):
    """Bucketing: shuffle the vectors, cut the shuffled list into buckets of ``bucket_size`` and return the bucket means.

    Averaging inside random buckets dilutes each Byzantine vector with honest ones and reduces the variance the
    aggregator that follows has to cope with; the output has ``ceil(n / bucket_size)`` vectors.

    Parameters
    ----------
    bucket_size : int
        Vectors per bucket (the last bucket may be smaller).
    feature_chunk_size : int, default 8192
        Coordinates per subtask on an actor pool.
    perm : iterable of int, optional
        A fixed permutation of ``range(n)`` to use instead of shuffling.
    rng : random.Random, optional
        Source of the shuffle; default: a fresh OS-seeded generator per instance, so two runs differ unless ``rng`` or
        ``perm`` is given.

    Notes
    -----
    The row map needs no Gram matrix.  In a fused device round the permutation is drawn on the host once per round
    (``MapCwPlan.refresh`` / the Gram plan's weight refresh) and uploaded as the mixing matrix.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.pre_aggregators import Bucketing
    >>> xs = [torch.tensor([float(i)]) for i in range(5)]
    >>> Bucketing(bucket_size=2, perm=[0, 1, 2, 3, 4]).pre_aggregate(xs)
    [tensor([0.5000]), tensor([2.5000]), tensor([4.])]
    """

    name = "pre-agg/bucketing"
    needs_gram = False

    def __init__(self, bucket_size: int, *, feature_chunk_size: int = 8192,
                 perm: Optional[Iterable[int]] = None, rng: Optional[random.Random] = None) -> None:
        if bucket_size < 1:
            raise ValueError("bucket_size must be >= 1")
        if feature_chunk_size <= 0:
            raise ValueError("feature_chunk_size must be > 0")
        self.bucket_size = int(bucket_size)
        self.feature_chunk_size = int(feature_chunk_size)
        self.perm = None if perm is None else [int(i) for i in perm]
        self.rng = rng or random.Random()

    def _resolve_order(self, n: int) -> List[int]:
        if self.perm is None:
            order = list(range(n))
            self.rng.shuffle(order)
            return order
        if len(self.perm) != n or sorted(self.perm) != list(range(n)):
            raise ValueError("perm must be a permutation of range(n)")
        return list(self.perm)

    def row_map(self, G, n):
        return nspace.bucket_matrix(n, self.bucket_size, self._resolve_order(n))


__all__ = ["Bucketing"]
