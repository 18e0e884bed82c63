"""Backend spec parsing (reference engine/actor/factory.py:14-67; the single copy used by
``configs.actor.set_actor`` and the node actors as well -- the reference has three)."""
from __future__ import annotations

from typing import Tuple, Union

from .base import ActorBackend


def _host_port(rest: str) -> Tuple[str, int]:
    host, port = rest.rsplit(":", 1)
    return host, int(port)


def resolve_backend(spec: Union[str, ActorBackend]) -> ActorBackend:
    """``"thread" | "process" | "gpu" | "tcp://host:port" | "ucx://host:port"`` or an instance."""
    if not isinstance(spec, str):
        return spec
    if spec == "thread":
        from .backends.thread import ThreadActorBackend

        return ThreadActorBackend()
    if spec == "process":
        from .backends.process import ProcessActorBackend

        return ProcessActorBackend()
    if spec == "gpu" or spec.startswith("gpu:"):
        from .backends.gpu import GPUActorBackend

        return GPUActorBackend(int(spec[4:]) if spec.startswith("gpu:") else None)
    if spec.startswith("tcp://"):
        from .backends.remote import RemoteActorBackend

        return RemoteActorBackend(*_host_port(spec[len("tcp://"):]))
    if spec.startswith("ucx://"):
        from .backends.gpu import UCXRemoteActorBackend

        return UCXRemoteActorBackend(*_host_port(spec[len("ucx://"):]))
    raise ValueError(f"Unknown actor backend spec: {spec!r}")


__all__ = ["resolve_backend"]
