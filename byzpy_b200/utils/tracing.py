"""Tracing / profiling helpers (the reference has none).

* ``nvtx_range(name)``  -- NVTX range when CUDA is present (visible in nsys / ncu), no-op otherwise.
* ``Tracer``            -- collects per-graph-node host wall time and (for CUDA work) device time
                           from CUDA events; schedulers pick it up from ``metadata["tracer"]``.
* ``cuda_time_ms(fn)``  -- device-timed call with warm-up, the recipe used for every number in
                           ``profiles/``.
"""
from __future__ import annotations

import contextlib
import time
from collections import defaultdict
from typing import Any, Callable, Dict, List, Optional

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    """Context manager pushing an NVTX range (visible in Nsight timelines); a no-op without CUDA."""
    pushed = False
    if torch.cuda.is_available():
        try:
            torch.cuda.nvtx.range_push(name)
            pushed = True
        except Exception:
            pushed = False
    try:
        yield
    finally:
        if pushed:
            torch.cuda.nvtx.range_pop()


class Tracer:
    """Lightweight span recorder: ``with tracer.span("name", **tags):`` measures host time and, on CUDA, device time
    between two events on the current stream (no synchronisation until ``finalize()``).  ``summary()`` aggregates calls,
    host and device milliseconds per span name; every span is also an NVTX range.
    """
    def __init__(self, cuda: Optional[bool] = None) -> None:
        self.cuda = torch.cuda.is_available() if cuda is None else cuda
        self.records: List[Dict[str, Any]] = []
        self._pending: List[tuple] = []

    @contextlib.contextmanager
    def span(self, name: str, **tags):
        t0 = time.perf_counter()
        ev = None
        if self.cuda:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        with nvtx_range(name):
            try:
                yield
            finally:
                rec = {"name": name, "host_ms": (time.perf_counter() - t0) * 1e3, **tags}
                if ev is not None:
                    ev[1].record()
                    self._pending.append((rec, ev))
                self.records.append(rec)

    def finalize(self) -> List[Dict[str, Any]]:
        if self._pending:
            torch.cuda.synchronize()
            for rec, (e0, e1) in self._pending:
                rec["device_ms"] = e0.elapsed_time(e1)
            self._pending.clear()
        return self.records

    def summary(self) -> Dict[str, Dict[str, float]]:
        agg: Dict[str, Dict[str, float]] = defaultdict(lambda: {"calls": 0, "host_ms": 0.0, "device_ms": 0.0})
        for r in self.finalize():
            a = agg[r["name"]]
            a["calls"] += 1
            a["host_ms"] += r.get("host_ms", 0.0)
            a["device_ms"] += r.get("device_ms", 0.0)
        return dict(agg)


def cuda_time_ms(fn: Callable[[], Any], *, warmup: int = 3, iters: int = 10) -> float:
    """Median device time of ``fn`` from CUDA events on the current stream."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


__all__ = ["nvtx_range", "Tracer", "cuda_time_ms"]
