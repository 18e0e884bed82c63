""""A little is enough" (Baruch et al.): ``mu + z * sigma`` per coordinate, with
``z = Phi^{-1}((N - s) / N)``, ``s = max(1, N // 2 + 1 - f)``, population std
(reference attacks/little.py:81-231)."""
from __future__ import annotations

from statistics import NormalDist
from typing import Optional

from .base import ColumnStatAttack


def _supporters_needed(N: int, f: int) -> int:
    if N <= 0:
        raise ValueError("N must be positive")
    return max(1, N // 2 + 1 - f)


def _ndtri(p: float) -> float:
    """Inverse standard-normal CDF (clamped away from 0 and 1)."""
    if not (0.0 <= p <= 1.0):
        raise ValueError("p must be in [0, 1]")
    tiny = 1e-12
    return NormalDist().inv_cdf(min(max(p, tiny), 1.0 - tiny))


class LittleAttack(ColumnStatAttack):
    """"A little is enough": shift every coordinate of the honest mean by ``z`` honest standard deviations.

    ``z`` is the largest shift that still leaves the malicious value among the majority a robust aggregator trusts:
    ``z = Phi^-1((N - s) / N)`` with ``s = max(1, N // 2 + 1 - f)`` supporters needed, ``Phi`` the standard normal CDF.
    The population standard deviation (divide by the number of honest gradients) is used.

    Parameters
    ----------
    f : int
        Number of Byzantine nodes.
    N : int, optional
        Total number of nodes; default: number of honest gradients given ``+ f``.
    chunk_size : int, default 8192
        Coordinates per subtask on an actor pool.

    Notes
    -----
    Needs ``honest_grads``.  ``z`` can be negative when fewer than half of the nodes have to be convinced.  In the
    fused device round mean and standard deviation of each coordinate come out of the aggregation kernel's own loads
    and the row is synthesised in registers.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.attacks import LittleAttack
    >>> atk = LittleAttack(f=2)
    >>> round(atk.z_value(8), 4)
    0.2533
    >>> atk.apply(honest_grads=[torch.tensor([0.0]), torch.tensor([2.0])] * 4)
    tensor([1.2533])
    """

    name = "little"
    max_subtasks_inflight = 0       # 0 / None: the pool-sized default window (value of the reference class)

    def __init__(self, f: int, N: Optional[int] = None, *, chunk_size: int = 8192) -> None:
        if f < 0:
            raise ValueError("f must be >= 0")
        if N is not None and N <= 0:
            raise ValueError("N must be positive")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.f = int(f)
        self.N = None if N is None else int(N)
        self.chunk_size = int(chunk_size)

    def z_value(self, n_honest: int) -> float:
        total = n_honest + self.f if self.N is None else self.N
        if total < self.f:
            raise ValueError(f"N must be >= f (got N={total}, f={self.f})")
        s = _supporters_needed(total, self.f)
        return _ndtri((total - s) / float(total))

    def _coeffs(self, n_honest: int):
        return 1.0, self.z_value(n_honest)


__all__ = ["LittleAttack"]
