"""``set_actor(spec)`` -> actor backend instance (reference configs/actor.py:10-30)."""
from __future__ import annotations

from typing import Union

from ..engine.actor.base import ActorBackend
from ..engine.actor.factory import resolve_backend


def set_actor(spec: Union[str, ActorBackend]) -> ActorBackend:
    """``"thread" | "process" | "gpu" | "tcp://host:port" | "ucx://host:port"``."""
    return resolve_backend(spec)


__all__ = ["set_actor"]
