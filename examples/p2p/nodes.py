"""P2P example nodes (counterpart of the reference's examples/p2p/nodes.py): SmallCNN honest node
with the P2P mixin, Empire Byzantine node."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseTrimmedMean
from byzpy_b200.attacks import EmpireAttack
from byzpy_b200.engine.node.mixin import P2PByzantineMixin, P2PHonestMixin
from byzpy_b200.models import SmallCNN
from byzpy_b200.parallel.arena import ParamArena
from byzpy_b200.utils.data import batch_source, mnist_like


class P2PHonestNode(P2PHonestMixin):
    def __init__(self, *, indices: Sequence[int], batch_size: int = 64, device: str = "cpu", f: int = 1, seed: int = 0):
        x, y = mnist_like(6000)
        idx = torch.as_tensor(list(indices))
        self._next = batch_source(x[idx], y[idx], batch_size, seed=seed)
        self.device = torch.device(device)
        torch.manual_seed(0)
        self.model = SmallCNN().to(self.device)
        self.arena = ParamArena(self.model)     # flat parameter / gradient buffers: the mixin's fast path
        self.criterion = nn.CrossEntropyLoss()
        self.optimizer = torch.optim.SGD(self.model.parameters(), lr=0.05)
        self.p2p_agg = CoordinateWiseTrimmedMean(f=f)
        self.p2p_pre = None

    def next_batch(self):
        x, y = self._next()
        return x.to(self.device), y.to(self.device)

    def dump_state_dict(self):
        return {k: v.detach().cpu() for k, v in self.model.state_dict().items()}


class P2PByzNode(P2PByzantineMixin):
    def __init__(self, *, device: str = "cpu", scale: float = -1.0):
        self.device = torch.device(device)
        self.attack = EmpireAttack(scale=scale)


__all__ = ["P2PHonestNode", "P2PByzNode"]
