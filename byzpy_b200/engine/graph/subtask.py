"""``SubTask`` -- the unit of work an operator hands to an ``ActorPool``.

Field-compatible with the reference dataclass (reference engine/graph/subtask.py:7-18): ``fn(*args,
**kwargs)`` plus scheduling hints.  ``affinity`` is a capability tag matched against the pool
workers (``"gpu"``, ``"cpu"``, or a worker tag ``"worker::<pool-name>-<index>"``); ``max_retries``
is honoured by ``ActorPool`` on worker failure.  On this framework a subtask that touches CUDA
tensors runs on a CUDA-stream worker with that stream current.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Any, Callable, Mapping, Optional, Sequence


@dataclass(frozen=True)
class SubTask:
    """A unit of work an operator hands to an :class:`~byzpy_b200.engine.graph.pool.ActorPool`.

    Parameters
    ----------
    fn : callable
        Runs on a worker as ``fn(*args, **kwargs)``.  It is shipped by value (cloudpickle) to process and remote
        workers, so it must not close over unpicklable state.
    args, kwargs :
        Arguments.  Tensors travel through shared memory to process workers and as CUDA-IPC handles to ``ucx://``
        workers; thread and GPU-stream workers receive the objects themselves.
    name : str, optional
        Label used in traces and error messages.
    affinity : str, optional
        Capability a worker must have: ``"cpu"``, ``"gpu"``, or one specific worker ``"worker::<pool-name>-<index>"``.
    max_retries : int, default 0
        How many times the pool re-runs the task on another worker after a worker failure.

    Examples
    --------
    >>> from byzpy_b200.engine.graph.subtask import SubTask
    >>> t = SubTask(fn=pow, args=(2, 10), name="pow")
    >>> t.run(), t.pinned_to("cpu").affinity
    (1024, 'cpu')
    """

    fn: Callable[..., Any]
    args: Sequence[Any] = field(default_factory=tuple)
    kwargs: Mapping[str, Any] = field(default_factory=dict)
    name: Optional[str] = None
    affinity: Optional[str] = None
    max_retries: int = 0

    def __post_init__(self) -> None:
        if not callable(self.fn):
            raise TypeError("SubTask.fn must be callable")
        if self.max_retries < 0:
            raise ValueError("SubTask.max_retries must be >= 0")

    # conveniences used by the schedulers / tests -------------------------------------------
    def run(self) -> Any:
        """Execute in the calling thread (what a worker does with the task)."""
        return self.fn(*self.args, **dict(self.kwargs))

    def pinned_to(self, affinity: Optional[str]) -> "SubTask":
        """Copy of this task with a different capability / worker tag."""
        return replace(self, affinity=affinity)

    @property
    def label(self) -> str:
        return self.name or getattr(self.fn, "__name__", "subtask")


__all__ = ["SubTask"]
