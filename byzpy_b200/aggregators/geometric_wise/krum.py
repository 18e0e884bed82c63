"""Krum and Multi-Krum (Blanchard et al.).

score_i = sum of the n-f-1 smallest squared distances from x_i to the others; Multi-Krum
averages the q best-scored vectors, Krum returns the single best (reference
aggregators/geometric_wise/krum.py:82-368).  Here: Gram pass (tcgen05 / CUDA-core),
single-CTA score+select solve, weighted-sum pass.
"""
from __future__ import annotations

import numpy as np

from ... import ops
from ...ops import nspace
from ..base import GramAggregator


class MultiKrum(GramAggregator):
    """Multi-Krum: average of the ``q`` gradients with the best Krum scores.

    The score of gradient ``i`` is the sum of its ``n - f - 1`` smallest squared distances to the other gradients; a
    gradient far from every honest cluster scores badly and is left out of the average.

    Parameters
    ----------
    f : int
        Upper bound on the number of Byzantine inputs; ``0 <= f < n - 1``.
    q : int
        Number of best-scored gradients averaged; ``1 <= q <= n - f``.
    chunk_size : int, default 32
        Rows per subtask on an actor pool (subtasks score row blocks against all rows).

    Notes
    -----
    Everything the operator needs is in the ``n x n`` Gram matrix: one pass over the data builds it
    (tcgen05 3xTF32 or exact fp32, ``ops.gram``), a single CTA turns it into ``q`` selected rows, one weighted-sum pass
    writes the result -- the gradients are read twice regardless of ``n``.  Ties in the score go to the lower index.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.geometric_wise import MultiKrum
    >>> honest = [torch.tensor([1.0, 1.0]) + 0.01 * i for i in range(4)]
    >>> MultiKrum(f=1, q=2).aggregate(honest + [torch.tensor([100.0, -100.0])])
    tensor([1.0150, 1.0150])
    """

    name = "multi-krum"
    shift_invariant = True       # distances only
    device_solve = True

    def __init__(self, f: int, q: int, *, chunk_size: int = 32) -> None:
        if f < 0:
            raise ValueError("f must be >= 0")
        if q < 1:
            raise ValueError("q must be >= 1")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.f = int(f)
        self.q = int(q)
        self.chunk_size = int(chunk_size)

    def _validate(self, n: int) -> None:
        if not (0 <= self.f < n - 1):
            raise ValueError(f"f must satisfy 0 <= f < n-1 (got n={n}, f={self.f})")
        if not (1 <= self.q <= n - self.f):
            raise ValueError(f"q must satisfy 1 <= q <= n - f (got n={n}, f={self.f}, q={self.q})")

    def _solve(self, G: np.ndarray, n: int) -> np.ndarray:
        return nspace.krum_weights(G, self.f, self.q)

    def _solve_device(self, G, n):
        from ...ops import nspace_cuda

        return nspace_cuda.krum_weights(G, self.f, self.q)

    def scores(self, gradients):
        """Krum scores of the inputs (diagnostic helper)."""
        from ..base import _kernel_rows, prepare_rows

        rows, _ = prepare_rows(gradients)
        G = ops.gram(_kernel_rows(rows), want64=True)
        return nspace.krum_scores(G.cpu().numpy(), self.f)


class Krum(MultiKrum):
    """Krum: the single gradient with the best Krum score (``MultiKrum`` with ``q = 1``).

    Parameters
    ----------
    f : int
        Upper bound on the number of Byzantine inputs; ``0 <= f < n - 1``.
    chunk_size : int, default 32
        Rows per subtask on an actor pool.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.geometric_wise import Krum
    >>> Krum(f=1).aggregate([torch.tensor([0.0]), torch.tensor([0.1]), torch.tensor([0.2]), torch.tensor([9.0])])
    tensor([0.1000])
    """

    name = "krum"

    def __init__(self, f: int, *, chunk_size: int = 32) -> None:
        super().__init__(f=f, q=1, chunk_size=chunk_size)


__all__ = ["MultiKrum", "Krum"]
