"""Attack base class (reference attacks/base.py:12-124).

An attack produces ONE malicious vector.  Which inputs it needs is declared by the class flags
``uses_base_grad`` / ``uses_model_batch`` / ``uses_honest_grads``; ``compute`` routes graph inputs
accordingly.  B200-native addition: ``fold()`` describes the attack as a *row fold* (a per-row
scale, or a synthesised ``a*mean + b*std`` row) so the device parameter server can apply it inside
the fused aggregation kernel without ever materialising the malicious vector.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict, List, Mapping, Optional, Sequence

import torch
import torch.nn as nn

from .. import ops
from ..aggregators.base import _Packed, _hold_packed, _kernel_rows, _release_packed, feature_chunks, finish, pool_in_process, pool_size_of, prepare_rows
from ..aggregators._chunking import select_adaptive_chunk_size
from ..engine.graph.operator import OpContext, Operator
from ..engine.graph.subtask import SubTask


class Attack(Operator, ABC):
    """Base class of attacks: an operator that produces one malicious vector for a Byzantine node to submit.

    Subclasses declare what they look at through three class flags and implement :meth:`apply`:

    * ``uses_base_grad`` -- the Byzantine node's own honest gradient (``base_grad``);
    * ``uses_honest_grads`` -- the gradients of the honest nodes this round (``honest_grads``), the omniscient setting;
    * ``uses_model_batch`` -- the node's model and a batch (``model``, ``x``, ``y``), for data-poisoning attacks.

    As a graph operator the attack reads exactly the inputs its flags name.  :meth:`fold` optionally describes the
    attack as something the fused device round can apply while it loads the rows (a per-row scale, an alias of another
    row, or a row synthesised from column statistics), in which case the malicious vector is never written to memory.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.attacks import Attack
    >>> class Zero(Attack):
    ...     name = "zero"
    ...     uses_base_grad = True
    ...     def apply(self, *, model=None, x=None, y=None, honest_grads=None, base_grad=None):
    ...         return torch.zeros_like(base_grad)
    >>> Zero().apply(base_grad=torch.ones(3))
    tensor([0., 0., 0.])
    """

    uses_base_grad: bool = False
    uses_model_batch: bool = False
    uses_honest_grads: bool = False

    name = "attack"
    supports_subtasks = False

    def compute(self, inputs: Mapping[str, Any], *, context: OpContext) -> Any:
        """Graph entry point: pick the inputs the attack's ``uses_*`` flags name out of ``inputs`` and call :meth:`apply`."""
        return self.apply(**self._collect_inputs(inputs))

    @abstractmethod
    def apply(self, *, model: Optional[nn.Module] = None, x: Optional[torch.Tensor] = None,
              y: Optional[torch.Tensor] = None, honest_grads: Optional[List[Any]] = None,
              base_grad: Optional[Any] = None) -> Any:
        """Return the malicious vector.  Only the keyword arguments the attack declared through its ``uses_*`` flags are
        required: ``base_grad``; ``honest_grads``; ``model`` with a batch ``x``, ``y``.
        """

    def fold(self, n_honest: int):
        """Row-fold description for the fused device round (None = must be materialised)."""
        return None

    def _collect_inputs(self, inputs: Mapping[str, Any]) -> Dict[str, Any]:
        picked: Dict[str, Any] = {}
        wanted = []
        if self.uses_model_batch:
            wanted += ["model", "x", "y"]
        if self.uses_honest_grads:
            wanted.append("honest_grads")
        if self.uses_base_grad:
            wanted.append("base_grad")
        for key in wanted:
            if key not in inputs:
                raise KeyError(f"Attack requires input {key!r}")
            picked[key] = inputs[key]
        return picked


def _colstat_chunk(packed: _Packed, start: int, end: int, a: float, b: float):
    rows = packed.slice(start, end)
    return start, ops.colstat(_kernel_rows(rows), a, b)


class ColumnStatAttack(Attack):
    """Shared machinery of the omniscient column-statistics attacks: the output is
    ``a * mean(honest) + b * std(honest)`` per coordinate (population std).

    Subclasses only supply ``_coeffs(n_honest) -> (a, b)``.  The direct call is one pass of the column-statistics
    kernel (``ops.colstat``), the pool path splits the coordinates, and the fused device round synthesises the row in
    registers from the loads the aggregation kernel makes anyway (``RowFold("virtual", a=a, b=b)``).

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.attacks.base import ColumnStatAttack
    >>> class MeanPlusStd(ColumnStatAttack):
    ...     name = "mean-plus-std"
    ...     def _coeffs(self, n_honest):
    ...         return 1.0, 1.0
    >>> MeanPlusStd().apply(honest_grads=[torch.tensor([0.0]), torch.tensor([2.0])])
    tensor([2.])
    """

    uses_honest_grads = True
    supports_subtasks = True
    chunk_size: int = 8192

    def _coeffs(self, n_honest: int):
        raise NotImplementedError

    def apply(self, *, model=None, x=None, y=None, honest_grads=None, base_grad=None):
        if not honest_grads:
            raise ValueError(f"{type(self).__name__} requires honest_grads.")
        rows, like = prepare_rows(honest_grads, "honest_grads")
        a, b = self._coeffs(len(rows))
        return finish(ops.colstat(_kernel_rows(rows), a, b), like)

    def fold(self, n_honest: int):
        from ..parallel.device_ps import RowFold

        a, b = self._coeffs(n_honest)
        return RowFold("virtual", a=a, b=b)

    def _subtask_feature_chunk(self, d: int, n_rows: int, context) -> int:
        """Coordinates per subtask; ``chunk_size`` counts coordinates for this family unless a subclass says
        otherwise."""
        return select_adaptive_chunk_size(d, self.chunk_size, pool_size=pool_size_of(context))

    def create_subtasks(self, inputs, *, context):
        grads = inputs.get("honest_grads")
        if not isinstance(grads, Sequence) or not grads:
            return []
        rows, _ = prepare_rows(grads, "honest_grads")
        d = rows[0].numel()
        a, b = self._coeffs(len(rows))
        chunk = self._subtask_feature_chunk(d, len(rows), context)
        packed = _Packed.pack(rows, in_process=pool_in_process(context))
        _hold_packed(self, inputs, packed)
        return [SubTask(fn=_colstat_chunk, args=(packed, s, e, a, b), name=f"{self.name}_chunk_{k}")
                for k, (s, e) in enumerate(feature_chunks(d, chunk))]

    def reduce_subtasks(self, partials, inputs, *, context):
        try:
            if not partials:
                return self.compute(inputs, context=context)
            _, like = prepare_rows(inputs["honest_grads"], "honest_grads")
            parts = sorted(partials, key=lambda p: p[0])
            vec = torch.cat([torch.as_tensor(p[1]).reshape(-1).to(like.device) for p in parts])
            return finish(vec, like)
        finally:
            _release_packed(self, inputs)


__all__ = ["Attack", "ColumnStatAttack"]
