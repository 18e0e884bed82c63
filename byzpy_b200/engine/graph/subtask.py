"""``SubTask``: a unit of work schedulable on a pool worker (CPU thread/process
actor or a CUDA-stream worker).  Field-compatible with the reference dataclass
(reference engine/graph/subtask.py:7-18)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Mapping, Optional, Sequence


@dataclass(frozen=True)
class SubTask:
    fn: Callable[..., Any]
    args: Sequence[Any] = field(default_factory=tuple)
    kwargs: Mapping[str, Any] = field(default_factory=dict)
    name: Optional[str] = None
    affinity: Optional[str] = None  # capability tag, e.g. "gpu" / "cpu" / "worker::<name>-<idx>"
    max_retries: int = 0


__all__ = ["SubTask"]
