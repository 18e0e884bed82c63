"""Pre-aggregator base (reference pre_aggregators/base.py:9-96): ``Sequence[vec] -> List[vec]``.

Every pre-aggregator in this library is a *linear row map* ``X' = W X`` whose matrix ``W``
(m x n) depends only on the Gram matrix of the inputs (norms, pairwise distances) or on nothing
at all (bucketing).  ``row_map`` exposes ``W`` so that a following Gram-family aggregator can be
composed in n-space -- ``G' = W G W^T``, final weights ``w_agg @ W`` -- and the pre-aggregated
vectors are never materialised.  ``pre_aggregate`` materialises them for API parity.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, List, Mapping, Optional, Sequence

import numpy as np
import torch

from .. import ops
from ..aggregators._chunking import select_adaptive_chunk_size
from ..aggregators.base import (_Packed, _gram_chunk, _hold_packed, _kernel_rows, _release_packed,
                                feature_chunks, finish, pool_in_process, pool_size_of, prepare_rows)
from ..engine.graph.operator import OpContext, Operator
from ..engine.graph.subtask import SubTask


class PreAggregator(Operator, ABC):
    """Base class of pre-aggregators: operators that map ``n`` vectors to ``m`` (usually better behaved) vectors.

    Subclasses implement :meth:`pre_aggregate`.  As an :class:`~byzpy_b200.engine.graph.operator.Operator` a
    pre-aggregator reads its input list from the key ``"vectors"`` of a computation graph, so it can be chained
    in front of an aggregator (``ParameterServer(..., pre_aggregator=...)``, ``GraphBuilder().input("vectors").apply(...)``).

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.pre_aggregators import PreAggregator
    >>> class Halve(PreAggregator):
    ...     name = "pre-agg/halve"
    ...     def pre_aggregate(self, xs):
    ...         return [0.5 * x for x in xs]
    >>> Halve().pre_aggregate([torch.tensor([2.0]), torch.tensor([4.0])])
    [tensor([1.]), tensor([2.])]
    """

    name = "pre_aggregator"
    input_key = "vectors"

    def compute(self, inputs: Mapping[str, Any], *, context: OpContext) -> Any:
        """Graph entry point: pre-aggregate the sequence found under ``inputs["vectors"]``."""
        if self.input_key not in inputs:
            raise KeyError(f"{self.name} expects input key {self.input_key!r}")
        xs = inputs[self.input_key]
        if not isinstance(xs, Sequence):
            raise TypeError(f"{self.name} expects a sequence at {self.input_key!r}")
        return self.pre_aggregate(xs)

    @abstractmethod
    def pre_aggregate(self, xs: Sequence[Any]) -> List[Any]:
        """Map the ``n`` input vectors to a list of (possibly fewer) vectors of the same shape, dtype and device."""


def _mix_chunk(packed: _Packed, start: int, end: int, W: torch.Tensor):
    rows = _kernel_rows(packed.slice(start, end))
    return start, ops.weighted_sum(rows, W.to(rows[0].device))


class LinearPreAggregator(PreAggregator):
    """Pre-aggregators expressible as a row map ``X' = W X`` with an ``m x n`` matrix that depends on the inputs only
    through their Gram matrix (or on nothing, as for bucketing).

    Subclasses implement :meth:`row_map` (host, fp64) and optionally :meth:`row_map_device`.  Exposing ``W`` is what
    lets a pre-aggregator compose with what follows without writing the intermediate vectors: a Gram-family
    aggregator runs on ``W G W^T`` and returns ``(w W) X``; a coordinate-wise aggregator after a diagonal map
    takes the scales as per-row factors inside its kernel.  :meth:`pre_aggregate` materialises ``W X`` (one
    multi-row weighted-sum pass) for callers that want the vectors.

    Examples
    --------
    >>> import numpy as np, torch
    >>> from byzpy_b200.pre_aggregators.base import LinearPreAggregator
    >>> class PairMeans(LinearPreAggregator):
    ...     name = "pre-agg/pair-means"
    ...     needs_gram = False
    ...     def row_map(self, G, n):
    ...         W = np.zeros((n // 2, n))
    ...         for k in range(n // 2):
    ...             W[k, 2 * k] = W[k, 2 * k + 1] = 0.5
    ...         return W
    >>> PairMeans().pre_aggregate([torch.tensor([0.0]), torch.tensor([2.0]), torch.tensor([4.0]), torch.tensor([8.0])])
    [tensor([1.]), tensor([6.])]
    """

    supports_subtasks = True
    max_subtasks_inflight = 0
    needs_gram: bool = True
    diagonal_map: bool = False      # W = diag(s): every output row reads exactly one input row
    feature_chunk_size: int = 8192

    def _validate(self, n: int) -> None:
        pass

    @abstractmethod
    def row_map(self, G: Optional[np.ndarray], n: int) -> np.ndarray:
        """(m, n) mixing matrix from the fp64 Gram matrix (``G`` is None when ``needs_gram`` is False)."""

    def row_map_device(self, G: torch.Tensor, n: int) -> Optional[torch.Tensor]:
        """The same matrix computed ON THE DEVICE from the device Gram (fp64, (m, n)), without a host
        round trip -- or None when the map has no device kernel (the caller then uses :meth:`row_map`)."""
        return None

    def _materialise(self, rows: List[torch.Tensor], W: np.ndarray, like: torch.Tensor) -> List[torch.Tensor]:
        krows = _kernel_rows(rows)
        Wt = torch.from_numpy(np.asarray(W, dtype=np.float32)).to(rows[0].device)
        if W.shape[0] == W.shape[1] and np.count_nonzero(W - np.diag(np.diagonal(W))) == 0:
            # diagonal map (clipping): n independent scaled copies
            diag = np.diagonal(W)
            return [finish(ops.scale_copy(krows[i], float(diag[i])), like) for i in range(len(rows))]
        Y = ops.weighted_sum(krows, Wt)
        if like.dim() == 1 and Y.dtype == like.dtype and Y.device == like.device:
            return list(Y.unbind(0))                 # already in the caller's shape / dtype / device
        return [finish(Y[i], like) for i in range(Y.shape[0])]

    def pre_aggregate(self, xs: Sequence[Any]) -> List[Any]:
        rows, like = prepare_rows(xs, "xs")
        n = len(rows)
        self._validate(n)
        G = None
        if self.needs_gram:
            Gd = ops.gram(_kernel_rows(rows), want64=True, diag_only=getattr(self, "gram_diag_only", False))
            if Gd.is_cuda:
                # sm_100a path: the n-space map is a single-CTA kernel on the device Gram and feeds the
                # weighted-sum pass directly -- no host synchronisation (reference nnm.py:82-97 also stays
                # on the input device; its clipping / ARC go through NumPy, clipping.py:53-62)
                Wd = self.row_map_device(Gd, n)
                if Wd is not None:
                    # diagonal maps (Clipping, ARC) read every input once with the 8-rows-per-pass kernel;
                    # dense maps (NNM) take the one-pass multi-row kernel
                    Y = ops.weighted_sum(_kernel_rows(rows), Wd.to(torch.float32),
                                         multi_impl="passes" if getattr(self, "diagonal_map", False) else "auto")
                    if like.dim() == 1 and Y.dtype == like.dtype:
                        return list(Y.unbind(0))
                    return [finish(Y[i], like) for i in range(Y.shape[0])]
            G = Gd.detach().cpu().numpy()
        return self._materialise(rows, self.row_map(G, n), like)

    # -- subtask path: split-K Gram over feature chunks, then one local materialisation ----
    # (maps that need no Gram -- Bucketing -- are materialised by the pool instead: every subtask emits
    #  the (m, chunk) slab of Y = W X for its feature chunk; the mixing matrix is drawn once, here)
    def create_subtasks(self, inputs, *, context):
        xs = inputs.get(self.input_key)
        if not isinstance(xs, Sequence) or not xs:
            return []
        rows, _ = prepare_rows(xs, "xs")
        self._validate(len(rows))
        d = rows[0].numel()
        if not self.needs_gram:
            W = torch.from_numpy(np.asarray(self.row_map(None, len(rows)), dtype=np.float32))
            chunk = select_adaptive_chunk_size(d, self.feature_chunk_size, pool_size=pool_size_of(context))
            packed = _Packed.pack(_kernel_rows(rows), in_process=pool_in_process(context))
            _hold_packed(self, inputs, packed)
            return [SubTask(fn=_mix_chunk, args=(packed, s, e, W), name=f"{self.name}_mix_{k}")
                    for k, (s, e) in enumerate(feature_chunks(d, chunk))]
        chunk = select_adaptive_chunk_size(d, self.feature_chunk_size, pool_size=pool_size_of(context))
        packed = _Packed.pack(_kernel_rows(rows), in_process=pool_in_process(context))
        _hold_packed(self, inputs, packed)
        return [SubTask(fn=_gram_chunk, args=(packed, s, e), name=f"{self.name}_gram_{k}")
                for k, (s, e) in enumerate(feature_chunks(d, chunk))]

    def reduce_subtasks(self, partials, inputs, *, context):
        try:
            if not partials:
                return self.compute(inputs, context=context)
            rows, like = prepare_rows(inputs[self.input_key], "xs")
            n = len(rows)
            if isinstance(partials[0], tuple):          # (offset, (m, chunk) slab) from _mix_chunk
                parts = sorted(partials, key=lambda p: p[0])
                Y = torch.cat([torch.as_tensor(p[1]).to(like.device) for p in parts], dim=1)
                return [finish(Y[i], like) for i in range(Y.shape[0])]
            G = np.zeros((n, n), dtype=np.float64)
            for p in partials:
                G += np.asarray(p, dtype=np.float64)
            return self._materialise(rows, self.row_map(G, n), like)
        finally:
            _release_packed(self, inputs)


__all__ = ["PreAggregator", "LinearPreAggregator"]
