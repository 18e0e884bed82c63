"""Label-flip attack: gradient of the loss on flipped labels (``y -> K-1-y`` or an explicit
mapping), scaled; leaves the model's grads zeroed (reference attacks/label_flip.py:35-91)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from ..parallel.arena import flatten_grads
from .base import Attack


def _flip(y: torch.Tensor, num_classes: Optional[int], mapping: Optional[Dict[int, int]]) -> torch.Tensor:
    if mapping is not None:
        out = y.clone()
        for src, dst in mapping.items():
            out[y == int(src)] = int(dst)
        return out
    return (int(num_classes) - 1) - y


class LabelFlipAttack(Attack):
    name = "label-flip"
    uses_model_batch = True

    def __init__(self, *, num_classes: Optional[int] = None, mapping: Optional[Dict[int, int]] = None,
                 loss_fn: Optional[nn.Module] = None, scale: float = 1.0) -> None:
        if mapping is None and num_classes is None:
            raise ValueError("Provide either `mapping` or `num_classes`.")
        self.num_classes = num_classes
        self.mapping = mapping
        self.loss_fn = loss_fn or nn.CrossEntropyLoss(reduction="mean")
        self.scale = float(scale)

    def apply(self, *, model=None, x=None, y=None, honest_grads=None, base_grad=None):
        if model is None or x is None or y is None:
            raise ValueError("LabelFlipAttack requires model, x, y.")
        y = y.to(dtype=torch.long, device=x.device)
        y_bad = _flip(y, self.num_classes, self.mapping)
        for p in model.parameters():
            if p.grad is not None:
                p.grad.zero_()
        model.train(True)
        loss = self.loss_fn(model(x), y_bad)
        loss.backward()
        vec = self.scale * flatten_grads(model).detach()
        for p in model.parameters():
            if p.grad is not None:
                p.grad.detach_()
                p.grad.zero_()
        return vec


__all__ = ["LabelFlipAttack"]
