"""SMEA (smallest maximum eigenvalue averaging): among all (n-f)-subsets pick the one whose
empirical covariance has the smallest top eigenvalue, computed on the centred m x m Gram
blocks (reference aggregators/geometric_wise/smea.py:63-88), batched over subsets."""
from __future__ import annotations

import numpy as np

from ...ops import nspace
from ..base import GramAggregator


class SMEA(GramAggregator):
    """Smallest Maximum Eigenvalue Averaging: mean of the ``n - f`` gradients whose covariance is the least stretched.

    For every subset of size ``n - f`` the largest eigenvalue of its empirical covariance is computed; the subset with
    the smallest one is averaged.  A group of colluding outliers inflates the variance along one direction, which this
    criterion sees even when every single distance looks harmless.

    Parameters
    ----------
    f : int
        Number of gradients left out; ``0 <= 2 f < n``.
    chunk_size : int, default 256
        Subsets scored per subtask on an actor pool.

    Notes
    -----
    The non-zero spectrum of a subset's covariance equals that of its centred ``m x m`` Gram block, so after one Gram
    pass the search is dense linear algebra on tiny matrices, batched over subsets (``numpy.linalg.eigvalsh`` on the
    host, power iteration in ``csrc/nspace.cu`` on the device).  The number of subsets grows combinatorially: meant for
    ``n`` up to a few tens.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.geometric_wise import SMEA
    >>> grads = [torch.tensor([0.0, 0.0]), torch.tensor([0.2, 0.0]), torch.tensor([0.1, 0.1]), torch.tensor([9.0, -9.0])]
    >>> SMEA(f=1).aggregate(grads)
    tensor([0.1000, 0.0333])
    """

    name = "smea"
    shift_invariant = True       # distances only

    def __init__(self, f: int, *, chunk_size: int = 256) -> None:
        if f < 0:
            raise ValueError("f must be >= 0")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.f = int(f)
        self.chunk_size = int(chunk_size)

    def _validate(self, n: int) -> None:
        if not (0 <= 2 * self.f < n):
            raise ValueError(f"2f must be < n (got n={n}, f={self.f})")

    def _solve(self, G: np.ndarray, n: int) -> np.ndarray:
        return nspace.smea_weights(G, self.f)

    # exhaustive subset search on the device (csrc/nspace.cu) when C(n, n-f) is small enough;
    # returns None otherwise and the host search above takes over
    device_solve = True

    def _solve_device(self, G, n):
        from ...ops import nspace_cuda

        return nspace_cuda.subset_weights(G, n, n - self.f, "smea")

    def _device_solve_feasible(self, n: int) -> bool:
        from ...ops import nspace_cuda

        return nspace_cuda.subset_search_feasible(n, n - self.f)


__all__ = ["SMEA"]
