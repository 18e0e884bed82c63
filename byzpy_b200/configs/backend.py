"""Array-backend selection (reference configs/backend.py:12-49).

The reference's ``get_backend()`` unconditionally returns a fresh torch backend, so
``set_backend`` / ``use_backend`` are inert there.  The API is kept; here the
selection is honoured: ``"torch"`` (default) or ``"numpy"``.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Iterator

from ..engine.backend.ndarray import get_array_backend

_current = "torch"


def set_backend(backend) -> None:
    """``backend``: ``"torch"`` / ``"pytorch"`` / ``"numpy"``, or a backend object (anything with a ``name``
    attribute naming one of them; reference configs/backend.py:12-29 takes the same two forms)."""
    global _current
    name = backend if isinstance(backend, str) else getattr(backend, "name", None)
    if not isinstance(name, str):
        raise TypeError("backend must be a string or implement the _Backend protocol")
    name = {"pytorch": "torch"}.get(name.lower(), name.lower())
    get_array_backend(name)  # validates
    _current = name


def get_backend():
    """The array backend object currently selected for this process."""
    return get_array_backend(_current)


@contextmanager
def use_backend(backend) -> Iterator[None]:
    """Context manager: select array backend ``backend`` (``"torch"`` / ``"numpy"``) inside the block."""
    global _current
    prev = _current
    set_backend(backend)
    try:
        yield
    finally:
        _current = prev


__all__ = ["set_backend", "get_backend", "use_backend"]
