"""Fall-of-empires / inner-product manipulation: ``scale * mean(honest_grads)``
(reference attacks/empire.py:23-187)."""
from __future__ import annotations

from .base import ColumnStatAttack


class EmpireAttack(ColumnStatAttack):
    name = "empire"

    def __init__(self, scale: float = -1.0, *, chunk_size: int = 8) -> None:
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.scale = float(scale)
        self.chunk_size = int(chunk_size)

    def _coeffs(self, n_honest: int):
        return self.scale, 0.0


__all__ = ["EmpireAttack"]
