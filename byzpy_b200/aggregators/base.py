"""Aggregator base classes.

``Aggregator`` keeps the reference contract (reference aggregators/base.py:11-103):
``aggregate(gradients) -> tensor`` plus the ``Operator`` interface with
``input_key = "gradients"``.  Two family bases carry the B200-native structure:

* :class:`CoordinateWiseAggregator` -- one streaming selection-network pass
  (``ops.cw_select``); shards over the feature dimension.
* :class:`GramAggregator` -- Gram pass, n-space solve, weighted-sum pass
  (``ops.gram`` -> ``ops.nspace`` -> ``ops.weighted_sum``); the Gram shards over the
  feature dimension (split-K) and the solve never touches the data.

Both expose ``fused_plan()`` so the device parameter server can fold the operator
into its fused cross-GPU kernel, and both are stateless across calls (re-entrant).
Inputs may be tensors, ndarrays, ``SharedTensorHandle`` s or handle dicts.
"""
from __future__ import annotations

import os
from abc import ABC, abstractmethod
from typing import Any, Iterable, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import ops
from ..engine.graph.operator import OpContext, Operator
from ..engine.graph.subtask import SubTask
from ..engine.storage.shared_store import (SharedTensorHandle, attach_cached, cleanup_tensor, materialize,
                                           register_rows)
from ._chunking import select_adaptive_chunk_size


class Aggregator(Operator, ABC):
    """Base class of every gradient aggregator: ``n`` gradients in, one robust estimate out.

    Subclasses implement :meth:`aggregate`.  Inputs are a sequence of same-shaped tensors (any floating dtype, CPU or
    CUDA) or array-likes the active backend understands; the result has the shape, dtype and device of the first
    input.  As an :class:`~byzpy_b200.engine.graph.operator.Operator` an aggregator reads its input list from
    the key ``"gradients"`` of a computation graph, and -- when ``supports_subtasks`` is set -- can split its work
    over an :class:`~byzpy_b200.engine.graph.pool.ActorPool` (``create_subtasks`` / ``reduce_subtasks``).
    :meth:`fused_plan` is the hook through which the device parameter server runs the aggregator inside its fused
    round instead of calling :meth:`aggregate`.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators import Aggregator
    >>> class Mean(Aggregator):
    ...     name = "mean"
    ...     def aggregate(self, gradients):
    ...         return torch.stack(list(gradients)).mean(0)
    >>> Mean().aggregate([torch.tensor([1.0, 2.0]), torch.tensor([3.0, 6.0])])
    tensor([2., 4.])
    """

    name = "aggregator"
    input_key = "gradients"

    def compute(self, inputs: Mapping[str, Any], *, context: OpContext) -> Any:
        """Graph entry point: aggregate the sequence found under ``inputs["gradients"]``."""
        if self.input_key not in inputs:
            raise KeyError(f"{self.name} expects input key {self.input_key!r}")
        grads = inputs[self.input_key]
        if not isinstance(grads, Sequence):
            raise TypeError(f"{self.name} expects a sequence at {self.input_key!r}")
        return self.aggregate(grads)

    @abstractmethod
    def aggregate(self, gradients: Sequence[Any]) -> Any:
        """Reduce the gradients to one tensor of the same shape, dtype and device as the first of them."""

    def fused_plan(self, n: int):
        """Plan object for :class:`byzpy_b200.parallel.device_ps.DeviceRound` (or None)."""
        return None


# --------------------------------------------------------------------------- helpers
def prepare_rows(gradients: Sequence[Any], what: str = "gradients") -> Tuple[List[torch.Tensor], torch.Tensor]:
    """Materialise inputs -> (flat row tensors, template tensor for shape/dtype/device)."""
    if gradients is None or len(gradients) == 0:
        raise ValueError(f"{what} must be a non-empty sequence")
    tensors = [materialize(g) for g in gradients]
    like = tensors[0]
    dev = like.device
    rows = []
    for t in tensors:
        if t.device != dev:
            t = t.to(dev)
        if not t.dtype.is_floating_point:
            t = t.to(torch.float32)
        rows.append(t if t.dim() == 1 else t.reshape(-1))
    d = rows[0].numel()
    for r in rows:
        if r.numel() != d:
            raise ValueError("all gradients must have the same number of elements")
    if not like.dtype.is_floating_point:
        like = like.to(torch.float32)
    return rows, like


def finish(vec: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """Reshape / cast a flat result to the template's shape, dtype and device."""
    return vec.reshape(like.shape).to(dtype=like.dtype, device=like.device)


def _kernel_rows(rows: List[torch.Tensor]) -> List[torch.Tensor]:
    """CUDA rows are fed to the kernels as fp32 (upcast copies for bf16/fp16/fp64)."""
    if rows[0].is_cuda:
        return [r if r.dtype == torch.float32 else r.float() for r in rows]
    return rows


def pool_size_of(context: Optional[OpContext]) -> int:
    """Number of pool workers the running scheduler announced in the operator context (0: no pool)."""
    meta = (context.metadata if context is not None else None) or {}
    return int(meta.get("pool_size") or 0)


def pool_in_process(context: Optional[OpContext]) -> bool:
    """Workers share this address space (thread / gpu pools): subtasks can take views of the rows."""
    meta = (context.metadata if context is not None else None) or {}
    return bool(meta.get("pool_in_process"))


class _Packed:
    """Rows packaged for subtasks: one host shm matrix, or in-process device rows."""

    __slots__ = ("handle", "rows", "aux")

    def __init__(self, handle: Optional[SharedTensorHandle], rows: Optional[List[torch.Tensor]]):
        self.handle = handle
        self.rows = rows
        self.aux = None         # coordinator-side only (never pickled): aux rows computed for this invocation

    @classmethod
    def pack(cls, rows: List[torch.Tensor], in_process: bool = False) -> "_Packed":
        if rows[0].is_cuda or in_process:
            return cls(None, rows)      # by reference: no POSIX shm round trip for in-process workers
        dtype = torch.float64 if rows[0].dtype == torch.float64 else torch.float32
        return cls(register_rows([r.to(dtype) for r in rows]), None)      # rows copied straight into the segment

    def slice(self, start: int, end: int) -> List[torch.Tensor]:
        if self.rows is not None:
            return [r[start:end] for r in self.rows]
        arr = attach_cached(self.handle)                 # mapping reused by this invocation's other subtasks
        chunk = torch.from_numpy(np.array(arr[:, start:end], copy=True))
        del arr
        return [chunk[i] for i in range(chunk.shape[0])]

    def release(self) -> None:
        if self.handle is not None:
            cleanup_tensor(self.handle)

    def __getstate__(self):
        return (self.handle, self.rows)

    def __setstate__(self, st):
        self.handle, self.rows = st
        self.aux = None


def feature_chunks(d: int, chunk: int) -> Iterable[Tuple[int, int]]:
    """``(start, end)`` pairs cutting ``range(d)`` into pieces of at most ``chunk`` coordinates."""
    for s in range(0, d, chunk):
        yield s, min(d, s + chunk)


# --------------------------------------------------------------- coordinate-wise family
def _cw_chunk(packed: _Packed, start: int, end: int, mode: int, f: int):
    rows = packed.slice(start, end)
    out = ops.cw_select(_kernel_rows(rows), mode, f)
    return start, out


class CoordinateWiseAggregator(Aggregator):
    """Shared machinery of the coordinate-wise family (median, trimmed mean, mean of medians).

    Each output coordinate depends only on that coordinate of the ``n`` inputs, so the work splits along the feature
    dimension without communication: on CUDA one selection-network launch over all coordinates
    (``ops.cw_select``); on an actor pool one subtask per feature chunk (``chunk_size`` coordinates, adapted to the
    pool size), concatenated by ``reduce_subtasks``; in the fused parameter-server round each GPU selects over the
    coordinate shard it owns (``CwPlan``).  Subclasses set ``_mode`` (which statistic) and ``_f`` (how many values are
    trimmed).

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
    >>> from byzpy_b200.aggregators.base import CoordinateWiseAggregator
    >>> isinstance(CoordinateWiseMedian(), CoordinateWiseAggregator)
    True
    """

    supports_subtasks = True
    max_subtasks_inflight = 0
    _mode: int = ops.MODE_MEDIAN
    chunk_size: int = 8192

    def _f(self, n: int) -> int:
        return 0

    def _validate(self, n: int) -> None:
        pass

    def aggregate(self, gradients: Sequence[Any]) -> Any:
        rows, like = prepare_rows(gradients)
        self._validate(len(rows))
        out = ops.cw_select(_kernel_rows(rows), self._mode, self._f(len(rows)))
        return finish(out, like)

    def fused_plan(self, n: int):
        from ..parallel.device_ps import CwPlan

        self._validate(n)
        return CwPlan(self._mode, self._f(n))

    def create_subtasks(self, inputs, *, context):
        grads = inputs.get(self.input_key)
        if not isinstance(grads, Sequence) or not grads:
            return []
        rows, _ = prepare_rows(grads)
        n, d = len(rows), rows[0].numel()
        self._validate(n)
        chunk = select_adaptive_chunk_size(d, self.chunk_size, pool_size=pool_size_of(context))
        packed = _Packed.pack(rows, in_process=pool_in_process(context))
        mode, f = self._mode, self._f(n)

        def gen():
            for k, (s, e) in enumerate(feature_chunks(d, chunk)):
                yield SubTask(fn=_cw_chunk, args=(packed, s, e, mode, f), name=f"{self.name}_chunk_{k}")

        _hold_packed(self, inputs, packed)
        return gen()

    def reduce_subtasks(self, partials, inputs, *, context):
        try:
            if not partials:
                return self.compute(inputs, context=context)
            _, like = prepare_rows(inputs[self.input_key])
            parts = sorted(partials, key=lambda p: p[0])
            vec = torch.cat([torch.as_tensor(p[1]).reshape(-1).to(like.device) for p in parts])
            return finish(vec, like)
        finally:
            _release_packed(self, inputs)


def _hold_packed(op, inputs, packed) -> None:
    """Keep the shm package alive until ``reduce_subtasks`` for THIS invocation (keyed by the
    identity of the per-run inputs mapping, so one operator instance stays re-entrant)."""
    live = op.__dict__.setdefault("_live_packages", {})
    stale = live.pop(id(inputs), None)      # an invocation that died before its reduce step: free its segment
    if stale is not None and stale is not packed:
        stale.release()
    live[id(inputs)] = packed


def _release_packed(op, inputs) -> None:
    packed = op.__dict__.get("_live_packages", {}).pop(id(inputs), None)
    if packed is not None:
        packed.release()


# --------------------------------------------------------------------------- Gram family
def _gram_chunk(packed: _Packed, start: int, end: int, with_median: bool = False):
    """Split-K partial Gram of one feature chunk.  ``with_median``: the coordinate-wise (lower) median of
    the chunk is computed here too, appended as the last Gram row, and returned with its offset -- the
    Weiszfeld / centered-clipping start point is then built by the pool, not by the coordinator."""
    rows = _kernel_rows(packed.slice(start, end))
    if with_median:
        med = ops.cw_median(rows).reshape(-1)
        G = ops.gram(rows + [med], want64=True)
        return start, G.detach().cpu().numpy(), med.detach()
    G = ops.gram(rows, want64=True)
    return G.detach().cpu().numpy()


class GramAggregator(Aggregator):
    """Shared machinery of the distance / norm based aggregators: Gram pass -> small solve -> weighted sum.

    Every aggregator of this family returns ``sum_i w_i x_i`` where the weights depend on the inputs only through
    their Gram matrix ``G = X X^T`` (norms, pairwise distances, spectra of sub-blocks).  The gradients are therefore
    read exactly twice whatever the algorithm iterates over: once to build the ``n x n`` matrix (``ops.gram``:
    tcgen05 tensor cores in 3xTF32 or exact fp32 CUDA cores, fp64 accumulation across tiles), once for the weighted
    sum (``ops.weighted_sum``).  Subclasses implement ``_solve`` (host, NumPy fp64) and optionally ``_solve_device``
    (one CTA on the device Gram, which keeps the fused round free of host round trips); ``_aux_rows`` appends
    extra rows to the Gram (a start point, a probe direction).

    ``shift_invariant`` marks solvers that only use distances; they can run on a centred Gram (``center``,
    ``BYZPY_GRAM_CENTER``) when the gradients share a component much larger than their differences.

    Examples
    --------
    >>> import numpy as np, torch
    >>> from byzpy_b200.aggregators.base import GramAggregator
    >>> class NearestToOrigin(GramAggregator):
    ...     name = "nearest-to-origin"
    ...     def _solve(self, G, n):
    ...         w = np.zeros(n)
    ...         w[int(np.argmin(np.diag(G)[:n]))] = 1.0      # the row with the smallest norm
    ...         return w
    >>> NearestToOrigin().aggregate([torch.tensor([3.0, 4.0]), torch.tensor([0.0, 1.0]), torch.tensor([2.0, 2.0])])
    tensor([0., 1.])
    """

    supports_subtasks = True
    max_subtasks_inflight = 0
    # ``chunk_size`` keeps the reference's constructor meaning (rows / pairs / combinations per
    # subtask) and only matters to subclasses that chunk a combinatorial search; the split-K Gram
    # subtasks chunk the FEATURE dimension and use ``gram_feature_chunk`` (adapted to the pool size):
    # a 32-feature chunk would turn a 1.2 M-parameter gradient into 37 500 subtasks.
    chunk_size: int = 8192
    gram_feature_chunk: int = 1 << 16
    _gram_chunk_elems: int = 1 << 16
    # Conditioning of the distance matrix.  ``D_ij = G_ii + G_jj - 2 G_ij`` from an fp32-product Gram loses the
    # distances once the gradients share a component much larger than their differences (late training:
    # |G| ~ d * common^2, round-off ~ 1e-7 |G|, true distances far below that) -- the reference's
    # ``flat @ flat.T`` has the same flaw.  Every solver marked ``shift_invariant`` only uses distances
    # (or affine combinations with coefficients summing to one), so it may run on the Gram of the rows
    # translated by ANY common vector; ``center = "median"`` translates by the coordinate-wise median (inside the
    # honest cluster whenever a majority is honest, finite whenever a majority of the rows is), ``"row0"`` by the
    # first row.  Off by default: it costs a median pass and a centred copy of the rows.  Instance attribute, or
    # ``BYZPY_GRAM_CENTER=median|row0`` for the process.  The output is still ``sum_i w_i x_i`` over the ORIGINAL rows.
    shift_invariant: bool = False
    center: Optional[str] = None

    # -- hooks ---------------------------------------------------------------------
    def _validate(self, n: int) -> None:
        pass

    def _aux_rows(self, rows: List[torch.Tensor]) -> List[torch.Tensor]:
        """Extra rows appended to the Gram (e.g. the median start point of Weiszfeld)."""
        return []

    @abstractmethod
    def _solve(self, G: np.ndarray, n: int) -> np.ndarray:
        """Host solve: fp64 Gram of (rows + aux rows) -> weights over (rows + aux rows)."""

    def _solve_device(self, G: torch.Tensor, n: int) -> Optional[torch.Tensor]:
        """Optional sync-free device solve (CUDA n-space kernels); None -> host solve."""
        return None

    # -- direct path ---------------------------------------------------------------
    def _weights(self, all_rows: List[torch.Tensor], n: int, G: Optional[torch.Tensor] = None) -> torch.Tensor:
        if G is None:
            G = ops.gram(all_rows, want64=True, diag_only=getattr(self, "gram_diag_only", False))
        if G.is_cuda:
            w = self._solve_device(G, n)
            if w is not None:
                return w
        w_np = self._solve(G.detach().cpu().numpy().astype(np.float64), n)
        return torch.from_numpy(np.asarray(w_np, dtype=np.float32)).to(all_rows[0].device)

    def _centering(self) -> Optional[str]:
        mode = self.center if self.center is not None else (os.environ.get("BYZPY_GRAM_CENTER") or None)
        if mode in (None, "", "off", "0") or not self._is_shift_invariant():
            return None
        if mode not in ("median", "row0"):
            raise ValueError("center must be None, 'median' or 'row0'")
        return mode

    def _is_shift_invariant(self) -> bool:
        return bool(self.shift_invariant)

    def _aggregate_centered(self, krows: List[torch.Tensor], n: int, like: torch.Tensor, mode: str) -> Any:
        aux = self._aux_rows(krows)
        c = ops.cw_median(krows) if mode == "median" else krows[0]
        c = torch.nan_to_num(c, nan=0.0, posinf=0.0, neginf=0.0)          # a non-finite centre would poison every row
        shifted = [r - c for r in krows] + [a - c for a in aux]
        w = self._weights(shifted, n)
        return finish(ops.weighted_sum(krows + aux, w.reshape(-1)), like)

    def aggregate(self, gradients: Sequence[Any]) -> Any:
        rows, like = prepare_rows(gradients)
        n = len(rows)
        self._validate(n)
        krows = _kernel_rows(rows)
        mode = self._centering()
        if mode is not None:
            return self._aggregate_centered(krows, n, like, mode)
        fused = ops.gram_with_median(krows, want64=True) if self._fused_aux() == ("median",) else None
        if fused is not None:
            # median start point + Gram matrix in ONE pass over the rows (csrc/gram.cu, AUX variant)
            G, med = fused
            all_rows = krows + [med]
            w = self._weights(all_rows, n, G=G)
        else:
            all_rows = krows + self._aux_rows(krows)
            w = self._weights(all_rows, n)
        out = ops.weighted_sum(all_rows, w.reshape(-1))
        return finish(out, like)

    # fused device round: auxiliary rows the solver expects after the real ones, and whether a
    # sync-free device solver exists (otherwise the round does one host round trip on G)
    fused_aux: Tuple[str, ...] = ()
    device_solve: bool = False

    def _fused_aux(self) -> Optional[Tuple[str, ...]]:
        """Aux rows as understood by DeviceRound, or None if this instance cannot be fused."""
        if type(self)._aux_rows is GramAggregator._aux_rows:
            return ()
        return None

    def fused_plan(self, n: int):
        from ..parallel.device_ps import GramPlan

        self._validate(n)
        aux = self._fused_aux()
        if aux is None:
            return None

        def solver(G: torch.Tensor) -> torch.Tensor:
            if self.device_solve and G.is_cuda:
                w = self._solve_device(G.double(), n)
                if w is not None:
                    return w
            w_np = self._solve(G.detach().double().cpu().numpy(), n)
            return torch.from_numpy(np.asarray(w_np, dtype=np.float32)).to(G.device)

        feasible = getattr(self, "_device_solve_feasible", lambda n_: True)(n)
        return GramPlan(solver, self.name, aux=tuple(aux), capturable=bool(self.device_solve) and feasible)

    # -- subtask path: split-K partial Grams -------------------------------------------
    def _gram_subtasks(self, krows: List[torch.Tensor], context) -> Tuple[_Packed, List[SubTask]]:
        """Feature-chunk subtasks over the real rows (+ aux rows).  A median aux row is produced by the
        subtasks themselves; any other aux rows are computed once here and remembered on the package."""
        with_median = self._fused_aux() == ("median",)
        aux = [] if with_median else self._aux_rows(krows)
        all_rows = krows + aux
        d = all_rows[0].numel()
        # allow_small_chunks: a gradient no longer than one configured chunk still gets split across the
        # workers (down to chunk / BYZPY_CHUNK_MAX_SHRINK features) instead of becoming a single subtask
        chunk = select_adaptive_chunk_size(d, max(int(self.gram_feature_chunk), 1),
                                           pool_size=pool_size_of(context), allow_small_chunks=True)
        packed = _Packed.pack(all_rows, in_process=pool_in_process(context))
        packed.aux = aux
        extra = (True,) if with_median else ()
        tasks = [SubTask(fn=_gram_chunk, args=(packed, s, e) + extra, name=f"{self.name}_gram_{k}")
                 for k, (s, e) in enumerate(feature_chunks(d, chunk))]
        return packed, tasks

    def create_subtasks(self, inputs, *, context):
        grads = inputs.get(self.input_key)
        if not isinstance(grads, Sequence) or not grads:
            return []
        rows, _ = prepare_rows(grads)
        self._validate(len(rows))
        krows = _kernel_rows(rows)
        packed, tasks = self._gram_subtasks(krows, context)
        _hold_packed(self, inputs, packed)
        return tasks

    def _finish_from_partials(self, partials, gradients, aux: Optional[List[torch.Tensor]] = None):
        rows, like = prepare_rows(gradients)
        n = len(rows)
        krows = _kernel_rows(rows)
        if partials and isinstance(partials[0], tuple):
            # (offset, partial Gram incl. the median row, median chunk): stitch the start point together
            parts = sorted(partials, key=lambda p: p[0])
            med = torch.cat([torch.as_tensor(p[2]).reshape(-1).to(krows[0].device) for p in parts])
            aux, mats = [med.to(krows[0].dtype)], [p[1] for p in parts]
        else:
            aux, mats = (self._aux_rows(krows) if aux is None else aux), partials
        all_rows = krows + list(aux)
        G = np.zeros((len(all_rows), len(all_rows)), dtype=np.float64)
        for p in mats:
            G += np.asarray(p, dtype=np.float64)
        w = self._weights(all_rows, n, torch.from_numpy(G))
        return finish(ops.weighted_sum(all_rows, w.reshape(-1)), like)

    def reduce_subtasks(self, partials, inputs, *, context):
        try:
            if not partials:
                return self.compute(inputs, context=context)
            packed = self.__dict__.get("_live_packages", {}).get(id(inputs))
            return self._finish_from_partials(partials, inputs[self.input_key],
                                              aux=None if packed is None else packed.aux)
        finally:
            _release_packed(self, inputs)

    async def run_barriered_subtasks(self, inputs, *, context, pool):
        """One barrier: fan out the split-K Gram, then solve + weighted sum locally."""
        grads = inputs.get(self.input_key)
        if not isinstance(grads, Sequence) or not grads:
            raise ValueError(f"{self.name} requires a non-empty gradient list.")
        rows, _ = prepare_rows(grads)
        self._validate(len(rows))
        krows = _kernel_rows(rows)
        packed, tasks = self._gram_subtasks(krows, context)
        try:
            partials = await self._run_subtasks(pool, tasks, self.max_subtasks_inflight, context)
            return self._finish_from_partials(partials, grads, aux=packed.aux)
        finally:
            packed.release()


__all__ = ["Aggregator", "CoordinateWiseAggregator", "GramAggregator", "prepare_rows", "finish"]
