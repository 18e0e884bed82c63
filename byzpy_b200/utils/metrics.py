"""Process-wide metrics registry: counters, gauges and histograms with a Prometheus text exposition.

The reference has no observability beyond ``print``.  Here the orchestrators, the actor pool and the
operator dispatcher report into one registry that is OFF by default -- an instrumented call site costs one attribute
read until :func:`enable` is called -- and can be read three ways: :func:`snapshot` (a dict, for tests and progress
bars), :func:`to_prometheus_text` (the text format scrapers understand), :func:`serve` (a tiny HTTP endpoint on a
daemon thread, standard library only).

    from byzpy_b200.utils import metrics
    metrics.enable()
    ...                                   # train
    print(metrics.snapshot()["byzpy_ps_rounds_total"])
    metrics.serve(port=9464)              # curl localhost:9464/metrics

Names reported by the framework:

``byzpy_ps_rounds_total``                rounds completed by ``ParameterServer`` (label ``path``: ``generic`` / ``device``)
``byzpy_ps_round_seconds``               histogram of host-side round durations (generic path)
``byzpy_ps_node_failures_total``         nodes skipped because they raised or timed out (label ``kind``)
``byzpy_p2p_rounds_total``               gossip rounds completed by ``DecentralizedPeerToPeer``
``byzpy_pool_subtasks_total``            subtasks executed by ``ActorPool`` workers
``byzpy_pool_subtask_retries_total``     subtask attempts that failed and were retried
``byzpy_operator_runs_total``            ``Operator.run`` calls that had a choice (labels ``op``, ``route``: ``pool`` / ``direct``)
``byzpy_node_handler_errors_total``      message handlers of decentralized nodes that raised (label ``type``)
"""
from __future__ import annotations

import bisect
import threading
import time
from contextlib import contextmanager
from typing import Dict, Iterator, List, Mapping, Optional, Sequence, Tuple

LabelKey = Tuple[Tuple[str, str], ...]
DEFAULT_BUCKETS = (0.0005, 0.001, 0.0025, 0.005, 0.01, 0.025, 0.05, 0.1, 0.25, 0.5, 1.0, 2.5, 5.0, 10.0)


def _key(labels: Optional[Mapping[str, object]]) -> LabelKey:
    return tuple(sorted((str(k), str(v)) for k, v in (labels or {}).items()))


class MetricsRegistry:
    """Thread-safe store of named series; one instance (:data:`REGISTRY`) serves the whole process."""

    def __init__(self) -> None:
        self._lock = threading.Lock()
        self.enabled = False
        self._counters: Dict[str, Dict[LabelKey, float]] = {}
        self._gauges: Dict[str, Dict[LabelKey, float]] = {}
        self._hists: Dict[str, Dict[LabelKey, List[float]]] = {}      # per series: bucket counts..., sum, count
        self._buckets: Dict[str, Tuple[float, ...]] = {}
        self._help: Dict[str, str] = {}

    # ------------------------------------------------------------------ writing
    def describe(self, name: str, text: str) -> None:
        """Attach a help string to ``name`` (shown in the exposition)."""
        self._help[name] = text

    def inc(self, name: str, amount: float = 1.0, labels: Optional[Mapping[str, object]] = None) -> None:
        """Add ``amount`` to counter ``name`` (created on first use)."""
        if not self.enabled:
            return
        k = _key(labels)
        with self._lock:
            series = self._counters.setdefault(name, {})
            series[k] = series.get(k, 0.0) + amount

    def set(self, name: str, value: float, labels: Optional[Mapping[str, object]] = None) -> None:
        """Set gauge ``name`` to ``value``."""
        if not self.enabled:
            return
        with self._lock:
            self._gauges.setdefault(name, {})[_key(labels)] = float(value)

    def observe(self, name: str, value: float, labels: Optional[Mapping[str, object]] = None,
                buckets: Sequence[float] = DEFAULT_BUCKETS) -> None:
        """Record one observation in histogram ``name`` (bucket bounds are fixed by the first observation)."""
        if not self.enabled:
            return
        k = _key(labels)
        with self._lock:
            bounds = self._buckets.setdefault(name, tuple(buckets))
            series = self._hists.setdefault(name, {})
            row = series.get(k)
            if row is None:
                row = series[k] = [0.0] * (len(bounds) + 3)         # buckets, +Inf, sum, count
            row[bisect.bisect_left(bounds, value)] += 1
            row[-2] += value
            row[-1] += 1

    @contextmanager
    def timer(self, name: str, labels: Optional[Mapping[str, object]] = None) -> Iterator[None]:
        """Context manager observing the elapsed wall time of its block in histogram ``name``."""
        if not self.enabled:
            yield
            return
        t0 = time.perf_counter()
        try:
            yield
        finally:
            self.observe(name, time.perf_counter() - t0, labels)

    def reset(self) -> None:
        """Drop every series (the enabled flag stays as it is)."""
        with self._lock:
            self._counters.clear()
            self._gauges.clear()
            self._hists.clear()
            self._buckets.clear()

    # ------------------------------------------------------------------ reading
    def snapshot(self) -> Dict[str, object]:
        """``{name: value}`` for unlabelled series, ``{name: {label string: value}}`` for labelled ones; histograms
        appear as ``{"count": ..., "sum": ..., "mean": ...}``."""
        def fold(series: Dict[LabelKey, object]):
            if list(series) == [()]:
                return series[()]
            return {",".join(f"{a}={b}" for a, b in k): v for k, v in series.items()}

        with self._lock:
            out: Dict[str, object] = {}
            for name, series in self._counters.items():
                out[name] = fold(dict(series))
            for name, series in self._gauges.items():
                out[name] = fold(dict(series))
            for name, series in self._hists.items():
                out[name] = fold({k: {"count": int(r[-1]), "sum": r[-2], "mean": (r[-2] / r[-1]) if r[-1] else 0.0}
                                  for k, r in series.items()})
            return out

    def to_prometheus_text(self) -> str:
        """The registry in the Prometheus text exposition format (version 0.0.4)."""
        def lab(k: LabelKey, extra: str = "") -> str:
            parts = [f'{a}="{b}"' for a, b in k] + ([extra] if extra else [])
            return "{" + ",".join(parts) + "}" if parts else ""

        lines: List[str] = []
        with self._lock:
            for kind, table in (("counter", self._counters), ("gauge", self._gauges)):
                for name in sorted(table):
                    if name in self._help:
                        lines.append(f"# HELP {name} {self._help[name]}")
                    lines.append(f"# TYPE {name} {kind}")
                    for k, v in sorted(table[name].items()):
                        lines.append(f"{name}{lab(k)} {v:g}")
            for name in sorted(self._hists):
                if name in self._help:
                    lines.append(f"# HELP {name} {self._help[name]}")
                lines.append(f"# TYPE {name} histogram")
                bounds = self._buckets[name]
                for k, row in sorted(self._hists[name].items()):
                    running = 0.0
                    for b, c in zip(bounds, row):
                        running += c
                        le = 'le="%g"' % b
                        lines.append(f"{name}_bucket{lab(k, le)} {running:g}")
                    inf = 'le="+Inf"'
                    lines.append(f"{name}_bucket{lab(k, inf)} {row[-1]:g}")
                    lines.append(f"{name}_sum{lab(k)} {row[-2]:g}")
                    lines.append(f"{name}_count{lab(k)} {row[-1]:g}")
        return "\n".join(lines) + "\n"


REGISTRY = MetricsRegistry()
inc = REGISTRY.inc
set_gauge = REGISTRY.set
observe = REGISTRY.observe
timer = REGISTRY.timer
snapshot = REGISTRY.snapshot
to_prometheus_text = REGISTRY.to_prometheus_text
reset = REGISTRY.reset


def enable(on: bool = True) -> None:
    """Switch collection on (or off); off is the default and makes every reporting call a no-op."""
    REGISTRY.enabled = bool(on)


def enabled() -> bool:
    """Whether collection is on."""
    return REGISTRY.enabled


_server = None


def serve(port: int = 9464, host: str = "127.0.0.1"):
    """Expose :func:`to_prometheus_text` at ``http://host:port/metrics`` from a daemon thread; returns the
    ``http.server`` object (``.server_address`` holds the bound port, ``.shutdown()`` stops it).  Turns collection on."""
    import http.server

    class Handler(http.server.BaseHTTPRequestHandler):
        def do_GET(self):  # noqa: N802 (http.server API)
            if self.path.rstrip("/") not in ("", "/metrics"):
                self.send_error(404)
                return
            body = to_prometheus_text().encode()
            self.send_response(200)
            self.send_header("Content-Type", "text/plain; version=0.0.4; charset=utf-8")
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def log_message(self, *args):          # keep the training output clean
            pass

    global _server
    enable()
    srv = http.server.ThreadingHTTPServer((host, int(port)), Handler)
    srv.daemon_threads = True
    threading.Thread(target=srv.serve_forever, name="byzpy-metrics", daemon=True).start()
    _server = srv
    return srv


__all__ = ["MetricsRegistry", "REGISTRY", "inc", "set_gauge", "observe", "timer", "snapshot", "to_prometheus_text",
           "reset", "enable", "enabled", "serve", "DEFAULT_BUCKETS"]
