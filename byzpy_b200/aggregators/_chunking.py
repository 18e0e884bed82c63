"""Adaptive chunk sizing for operator subtasks.

Same contract and environment overrides as the reference helper (reference
aggregators/_chunking.py:30-72): keep roughly ``min_chunks_per_worker`` chunks per
pool worker in flight without shrinking a configured chunk by more than
``max_shrink_factor``.

  BYZPY_CHUNK_MIN_PER_WORKER   int >= 1, overrides ``min_chunks_per_worker``
  BYZPY_CHUNK_MAX_SHRINK       int >= 1, overrides ``max_shrink_factor``
  BYZPY_CHUNK_TARGET_FACTOR    float > 0, scales the chunks-per-worker target
"""
from __future__ import annotations

import os
from typing import Optional


def _env_int(name: str) -> Optional[int]:
    raw = os.environ.get(name)
    if raw is None:
        return None
    try:
        return max(1, int(raw, 10))
    except ValueError:
        return None


def _env_float(name: str) -> Optional[float]:
    raw = os.environ.get(name)
    if raw is None:
        return None
    try:
        return max(0.0, float(raw))
    except ValueError:
        return None


def _ceil_div(a: int, b: int) -> int:
    return -(-a // b)


def select_adaptive_chunk_size(total_items: int, configured_chunk: int, *,
                               pool_size: Optional[int] = None, min_chunks_per_worker: int = 4,
                               max_shrink_factor: int = 8, allow_small_chunks: bool = False) -> int:
    """Subtask granularity for ``total_items`` items on a pool of ``pool_size`` workers.

    Starts from the operator's configured chunk and shrinks it (by at most ``max_shrink_factor``) until every worker has
    about ``min_chunks_per_worker`` chunks to steal from; small inputs keep the configured chunk unless
    ``allow_small_chunks``.  ``BYZPY_CHUNK_MIN_PER_WORKER`` / ``BYZPY_CHUNK_MAX_SHRINK`` / ``BYZPY_CHUNK_TARGET_FACTOR`` tune it
    (same names as the reference, aggregators/_chunking.py).
    """
    if total_items <= 0:
        return 0
    base = min(max(1, int(configured_chunk)), total_items)
    if not pool_size or pool_size <= 1:
        return base
    per_worker = max(1, _env_int("BYZPY_CHUNK_MIN_PER_WORKER") or int(min_chunks_per_worker))
    shrink = max(1, _env_int("BYZPY_CHUNK_MAX_SHRINK") or int(max_shrink_factor))
    factor = _env_float("BYZPY_CHUNK_TARGET_FACTOR")
    if not factor:
        factor = 1.0
    # enough work per worker already: keep the configured granularity
    if not allow_small_chunks and total_items <= max(1, int(configured_chunk)) * pool_size:
        return base
    have = _ceil_div(total_items, base)
    want = max(have, int(-(-per_worker * pool_size * factor // 1)))
    tuned = max(max(1, base // shrink), _ceil_div(total_items, want))
    return min(base, tuned, total_items)


__all__ = ["select_adaptive_chunk_size"]
