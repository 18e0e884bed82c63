"""Parsing of the common ``--pool-workers`` / ``--pool-backend`` benchmark flags (reference
benchmarks/pytorch/_worker_args.py:6-38): a comma separated list of worker counts and a backend spec
which may itself be a comma separated mix, e.g. ``thread,process,tcp://host:29000``."""
from __future__ import annotations

import argparse
from typing import List, Sequence

from byzpy_b200.engine.graph.pool import ActorPoolConfig


DEFAULT_WORKER_COUNTS = (2, 4, 6)


def parse_worker_counts(spec: str) -> List[int]:
    """``"2,4,6"`` / ``"2 4 6"`` -> [2, 4, 6]."""
    try:
        out = [int(x) for x in str(spec).replace(",", " ").split()]
    except ValueError:
        raise argparse.ArgumentTypeError(f"bad worker count in {spec!r}") from None
    if not out or any(k < 1 for k in out):
        raise argparse.ArgumentTypeError(f"bad --pool-workers {spec!r} (example: 2,4,6)")
    return out


def coerce_worker_counts(value) -> List[int]:
    """A worker specification as it may sit in a namespace -- a string, one int, or a sequence of either -- as a
    list of ints."""
    if isinstance(value, str):
        return parse_worker_counts(value)
    if isinstance(value, int):
        return [value]
    if isinstance(value, Sequence):
        out: List[int] = []
        for v in value:
            if not isinstance(v, (int, str)):
                raise TypeError(f"unsupported worker count entry {v!r}")
            out.extend(coerce_worker_counts(v))
        if not out:
            raise ValueError("empty worker count list")
        return out
    raise TypeError(f"cannot read worker counts from {value!r}")


def pool_configs(backend_spec: str, count: int) -> List[ActorPoolConfig]:
    """``count`` workers spread round-robin over the backends listed in ``backend_spec``."""
    backends = [b.strip() for b in str(backend_spec).split(",") if b.strip()] or ["thread"]
    per = {b: 0 for b in backends}
    for i in range(count):
        per[backends[i % len(backends)]] += 1
    return [ActorPoolConfig(backend=b, count=k) for b, k in per.items() if k > 0]
