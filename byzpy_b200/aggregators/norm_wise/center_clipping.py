"""Centered clipping (Karimireddy et al.): ``v <- v + (1/n) sum_i clip(x_i - v, c_tau)`` for M
rounds (reference aggregators/norm_wise/center_clipping.py:131-156).  Like Weiszfeld, the
iterate stays in the span of the rows and the start point, so the M rounds run on the Gram
matrix; the data is read twice in total instead of M times."""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from ... import ops
from ...ops import nspace
from ..base import GramAggregator


class CenteredClipping(GramAggregator):
    """Centered clipping: ``v <- v + mean_i clip(x_i - v, c_tau)`` repeated ``M`` times.

    Every gradient pulls the running estimate towards itself, but by at most ``c_tau`` per round, so a far-away input
    has bounded influence.

    Parameters
    ----------
    c_tau : float
        Clipping radius (``>= 0``).
    M : int, default 10
        Number of rounds.
    eps : float, default 1e-12
        Lower clamp on a norm before dividing by it.
    init : {"mean", "median", "zero"}, default "mean"
        Start point of ``v``.
    chunk_size : int, default 32
        Rows per subtask; on an actor pool every round is one barriered round of subtasks.

    Notes
    -----
    As with the geometric median, ``v`` stays an affine combination of the inputs and the start point, so the ``M``
    rounds run on coefficients against the Gram matrix: two passes over the gradients in total.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.norm_wise import CenteredClipping
    >>> grads = [torch.tensor([1.0]), torch.tensor([1.1]), torch.tensor([0.9]), torch.tensor([1000.0])]
    >>> out = CenteredClipping(c_tau=0.5, M=20, init="median").aggregate(grads)
    >>> bool(out.item() < 5.0)
    True
    """

    name = "centered-clipping"
    supports_barriered_subtasks = True
    device_solve = True

    def _fused_aux(self):
        return ("median",) if self.init == "median" else ()

    def _is_shift_invariant(self) -> bool:
        # v <- v + (1/n) sum clip(x_i - v): the coefficients of v over the rows keep summing to one when they
        # start that way (mean / median start); the zero start is an absolute point, not an affine combination
        return self.init != "zero"

    def __init__(self, *, c_tau: float, M: int = 10, eps: float = 1e-12, init: str = "mean",
                 chunk_size: int = 32) -> None:
        if c_tau < 0:
            raise ValueError("c_tau must be >= 0")
        if M < 0:
            raise ValueError("M must be >= 0")
        if eps <= 0:
            raise ValueError("eps must be > 0")
        if init not in ("mean", "median", "zero"):
            raise ValueError("init must be 'mean', 'median' or 'zero'")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.c_tau, self.M, self.eps, self.init = float(c_tau), int(M), float(eps), init
        self.chunk_size = int(chunk_size)

    def _aux_rows(self, rows: List[torch.Tensor]) -> List[torch.Tensor]:
        if self.init == "median":
            return [ops.cw_median(rows)]  # lower median, as the reference's direct path
        return []

    def _start(self, n: int) -> np.ndarray:
        if self.init == "median":
            a0 = np.zeros(n + 1)
            a0[n] = 1.0
            return a0
        if self.init == "mean":
            return np.full(n, 1.0 / n)
        return np.zeros(n)

    def _solve(self, G: np.ndarray, n: int) -> np.ndarray:
        return nspace.centered_clip_coeffs(G, n, self._start(n), c_tau=self.c_tau, M=self.M,
                                           eps=self.eps)

    def _solve_device(self, G, n):
        from ...ops import nspace_cuda

        return nspace_cuda.centered_clip_coeffs(G, n, self._start(n), c_tau=self.c_tau, M=self.M,
                                                eps=self.eps)


__all__ = ["CenteredClipping"]
