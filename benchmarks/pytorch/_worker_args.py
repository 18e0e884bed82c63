"""Parsing of the common ``--pool-workers`` / ``--pool-backend`` benchmark flags (reference
benchmarks/pytorch/_worker_args.py:6-38): a comma separated list of worker counts and a backend spec
which may itself be a comma separated mix, e.g. ``thread,process,tcp://host:29000``."""
from __future__ import annotations

from typing import List

from byzpy_b200.engine.graph.pool import ActorPoolConfig


def parse_worker_counts(spec: str) -> List[int]:
    out = [int(x) for x in str(spec).split(",") if x.strip()]
    if not out or any(k < 1 for k in out):
        raise ValueError(f"bad --pool-workers {spec!r}")
    return out


def pool_configs(backend_spec: str, count: int) -> List[ActorPoolConfig]:
    """``count`` workers spread round-robin over the backends listed in ``backend_spec``."""
    backends = [b.strip() for b in str(backend_spec).split(",") if b.strip()] or ["thread"]
    per = {b: 0 for b in backends}
    for i in range(count):
        per[backends[i % len(backends)]] += 1
    return [ActorPoolConfig(backend=b, count=k) for b, k in per.items() if k > 0]
