"""Label-flip (data-poisoning) attack: the Byzantine node trains on corrupted targets and submits
the resulting gradient (reference attacks/label_flip.py:35-91).

Targets are remapped either with an explicit ``mapping`` {class -> class} or, given
``num_classes`` = K, by the involution ``y -> K - 1 - y``.  The gradient is returned flat, in
``model.parameters()`` order, multiplied by ``scale``; the model's ``.grad`` buffers are left zeroed so
the node's own optimizer state is not polluted.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from ..parallel.arena import flatten_grads
from .base import Attack


class LabelFlipAttack(Attack):
    """Label flip (data poisoning): the gradient of the loss on a batch whose targets were remapped.

    Parameters
    ----------
    num_classes : int, optional
        With ``K`` classes and no ``mapping`` the targets are flipped as ``y -> K - 1 - y``.
    mapping : dict of int to int, optional
        Explicit remapping of classes; classes not named keep their label.  One of ``num_classes`` / ``mapping`` is
        required.
    loss_fn : torch.nn.Module, optional
        Loss applied to ``model(x)`` and the corrupted targets; default mean cross entropy.
    scale : float, default 1.0
        Multiplier of the submitted gradient.

    Notes
    -----
    Needs ``model``, ``x``, ``y``.  The gradient is returned flat in ``model.parameters()`` order and the model's
    ``.grad`` buffers are cleared afterwards, so the node's optimizer state is not touched.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.attacks import LabelFlipAttack
    >>> atk = LabelFlipAttack(num_classes=10)
    >>> atk.corrupt(torch.tensor([0, 3, 9]))
    tensor([9, 6, 0])
    >>> model = torch.nn.Linear(4, 10)
    >>> atk.apply(model=model, x=torch.randn(2, 4), y=torch.tensor([1, 2])).shape
    torch.Size([50])
    """

    name = "label-flip"
    uses_model_batch = True

    def __init__(self, *, num_classes: Optional[int] = None, mapping: Optional[Dict[int, int]] = None,
                 loss_fn: Optional[nn.Module] = None, scale: float = 1.0) -> None:
        if num_classes is None and mapping is None:
            raise ValueError("Provide either `mapping` or `num_classes`.")
        self.num_classes, self.mapping = num_classes, mapping
        self.scale = float(scale)
        self.loss_fn = loss_fn if loss_fn is not None else nn.CrossEntropyLoss(reduction="mean")

    # -- target corruption ----------------------------------------------------------------------
    def corrupt(self, y: torch.Tensor) -> torch.Tensor:
        if self.mapping is None:
            return (int(self.num_classes) - 1) - y
        lut = torch.arange(int(max(int(y.max().item()) if y.numel() else 0,
                                   max(self.mapping), max(self.mapping.values()))) + 1, device=y.device)
        for src, dst in self.mapping.items():
            lut[int(src)] = int(dst)
        return lut[y]

    @staticmethod
    def _zero_grads(model: nn.Module) -> None:
        for p in model.parameters():
            if p.grad is not None:
                p.grad.detach_()
                p.grad.zero_()

    def apply(self, *, model=None, x=None, y=None, honest_grads=None, base_grad=None):
        if model is None or x is None or y is None:
            raise ValueError("LabelFlipAttack requires model, x, y.")
        poisoned = self.corrupt(y.to(device=x.device, dtype=torch.long))
        self._zero_grads(model)
        model.train(True)
        self.loss_fn(model(x), poisoned).backward()
        out = flatten_grads(model).detach() * self.scale
        self._zero_grads(model)
        return out


__all__ = ["LabelFlipAttack"]
