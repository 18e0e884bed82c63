"""Host shared-memory tensor store (CPU plumbing path).

API-compatible with the reference store (reference
engine/storage/shared_store.py:11-54): ``register_tensor`` copies an array into a
named POSIX shm segment and returns a picklable ``SharedTensorHandle``;
``open_tensor`` maps it as a NumPy view; ``cleanup_tensor`` unlinks it.

On the B200 path this store is bypassed entirely: device tensors are shared as
``DeviceTensorHandle`` s (CUDA-IPC addresses, see ``byzpy_b200.parallel.symmetric``)
and never bounce through the host.  Unlike the reference, ``handle.dtype`` being
a string is handled everywhere (``materialize`` below), so aggregators accept
handles directly (fixes the reference's ``ParameterServer`` + median TypeError,
SURVEY 0.4).
"""
from __future__ import annotations

from collections import OrderedDict
from contextlib import contextmanager
from dataclasses import dataclass
from multiprocessing import shared_memory
from typing import Any, Iterator, Tuple, Union

import numpy as np
import torch


@dataclass(frozen=True)
class SharedTensorHandle:
    """Name, shape and dtype of an array living in a POSIX shared-memory segment.

    Small and picklable: this is what crosses a process boundary instead of the data.  Operators accept handles (or the
    equivalent ``{"name", "shape", "dtype"}`` dict) wherever they accept tensors.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.engine.storage.shared_store import cleanup_tensor, open_tensor, register_tensor
    >>> h = register_tensor(torch.arange(6.0).reshape(2, 3))
    >>> with open_tensor(h) as view:
    ...     print(h.shape, h.dtype, float(view.sum()))
    (2, 3) float32 15.0
    >>> cleanup_tensor(h)
    """

    name: str
    shape: Tuple[int, ...]
    dtype: str


SharedHandleLike = Union[SharedTensorHandle, dict]


def _coerce(handle: SharedHandleLike) -> SharedTensorHandle:
    if isinstance(handle, SharedTensorHandle):
        return handle
    if isinstance(handle, dict):
        return SharedTensorHandle(name=handle["name"], shape=tuple(handle["shape"]),
                                  dtype=str(handle["dtype"]))
    raise TypeError(f"not a shared tensor handle: {type(handle)!r}")


def is_handle(obj: Any) -> bool:
    """True for a :class:`SharedTensorHandle` or a dict with its three keys."""
    return isinstance(obj, SharedTensorHandle) or (
        isinstance(obj, dict) and {"name", "shape", "dtype"} <= set(obj.keys()))


def register_tensor(array: Any) -> SharedTensorHandle:
    """Copy a tensor / array into a fresh shared-memory segment and return its handle.  The caller owns the
    segment: :func:`cleanup_tensor` unlinks it."""
    if isinstance(array, torch.Tensor):
        array = array.detach().cpu().numpy()
    arr = np.ascontiguousarray(array)
    seg = shared_memory.SharedMemory(create=True, size=max(1, arr.nbytes))
    try:
        np.ndarray(arr.shape, dtype=arr.dtype, buffer=seg.buf)[...] = arr
    finally:
        seg.close()
    return SharedTensorHandle(name=seg.name, shape=tuple(arr.shape), dtype=str(arr.dtype))


def register_rows(rows) -> SharedTensorHandle:
    """``n`` equally long 1-D tensors -> one ``(n, d)`` shared segment, each row copied straight into the
    mapping (one pass over the data; ``register_tensor(torch.stack(rows))`` makes two and a temporary)."""
    rows = list(rows)
    n, d = len(rows), int(rows[0].numel())
    dtype = rows[0].dtype
    itemsize = torch.empty((), dtype=dtype).element_size()
    seg = shared_memory.SharedMemory(create=True, size=max(1, n * d * itemsize))
    try:
        if n * d:
            view = torch.frombuffer(seg.buf, dtype=dtype, count=n * d).view(n, d)
            for i, r in enumerate(rows):       # row-wise copy_: 5 ms for 17 MB; torch.stack(out=view) takes 150 ms
                view[i].copy_(r.detach().reshape(-1))
            del view
    finally:
        seg.close()
    return SharedTensorHandle(name=seg.name, shape=(n, d), dtype=str(torch.empty((), dtype=dtype).numpy().dtype))


_ATTACHED: "OrderedDict[str, shared_memory.SharedMemory]" = OrderedDict()
_ATTACH_LIMIT = 4


def attach_cached(handle: SharedHandleLike) -> np.ndarray:
    """Map a segment and keep the mapping for the next few calls: the subtasks of one operator invocation
    all read the same segment, and ``shm_open`` + ``mmap`` + page-table population per subtask is most of a
    small subtask's cost.  At most ``_ATTACH_LIMIT`` mappings stay open per process (oldest closed first);
    the creator's ``cleanup_tensor`` still unlinks the name, the memory goes when the last mapping closes."""
    h = _coerce(handle)
    seg = _ATTACHED.get(h.name)
    if seg is None:
        seg = shared_memory.SharedMemory(name=h.name)
        _ATTACHED[h.name] = seg
        while len(_ATTACHED) > _ATTACH_LIMIT:
            _, old = _ATTACHED.popitem(last=False)
            try:
                old.close()
            except BufferError:            # a view is still alive somewhere: let the GC close it
                pass
    else:
        _ATTACHED.move_to_end(h.name)
    return np.ndarray(h.shape, dtype=np.dtype(h.dtype), buffer=seg.buf)


@contextmanager
def open_tensor(handle: SharedHandleLike) -> Iterator[np.ndarray]:
    """Context manager mapping a segment as a NumPy array (a view, not a copy; do not keep it past the block)."""
    h = _coerce(handle)
    seg = shared_memory.SharedMemory(name=h.name)
    try:
        yield np.ndarray(h.shape, dtype=np.dtype(h.dtype), buffer=seg.buf)
    finally:
        seg.close()


def cleanup_tensor(handle: SharedHandleLike) -> None:
    """Unlink the segment (idempotent); the memory is freed once the last mapping is closed."""
    h = _coerce(handle)
    try:
        seg = shared_memory.SharedMemory(name=h.name)
    except FileNotFoundError:
        return
    try:
        seg.unlink()
    finally:
        seg.close()


def materialize(obj: Any) -> torch.Tensor:
    """Tensor / ndarray / handle / handle-dict / sequence -> ``torch.Tensor`` (copies out of shm)."""
    if isinstance(obj, torch.Tensor):
        return obj
    if is_handle(obj):
        with open_tensor(obj) as arr:
            return torch.from_numpy(np.array(arr, copy=True))
    if isinstance(obj, np.ndarray):
        return torch.from_numpy(obj)
    return torch.as_tensor(obj)


__all__ = ["SharedTensorHandle", "register_tensor", "register_rows", "attach_cached", "open_tensor",
           "cleanup_tensor", "is_handle", "materialize"]
