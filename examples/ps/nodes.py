"""Canonical user node classes for the parameter-server examples (counterpart of the reference's
examples/ps/nodes.py): an honest SmallCNN worker and an Empire-attack Byzantine worker built on the
``Distributed*Node`` bases, so aggregation / gradient / attack run through per-node pipelines."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
from byzpy_b200.attacks import EmpireAttack
from byzpy_b200.engine.graph.pool import ActorPoolConfig
from byzpy_b200.engine.node.distributed import DistributedByzantineNode, DistributedHonestNode
from byzpy_b200.models import SmallCNN
from byzpy_b200.parallel.arena import flatten_grads, write_vector_to_grads_
from byzpy_b200.utils.data import batch_source, mnist_like


def select_pool_backend(spec: str) -> str:
    if spec == "process" or spec.startswith("tcp://"):
        return "thread"
    if spec.startswith("ucx://"):
        return "gpu"
    return spec


class DistributedPSHonestNode(DistributedHonestNode):
    def __init__(self, *, indices: Sequence[int], batch_size: int = 64, lr: float = 0.05,
                 momentum: float = 0.9, device: str = "cpu", pool_backend: str = "thread", seed: int = 0):
        super().__init__(actor_pool=[ActorPoolConfig(backend=pool_backend, count=1, name="worker")],
                         aggregator=CoordinateWiseMedian(), name=f"honest-{pool_backend}")
        x, y = mnist_like(6000)
        idx = torch.as_tensor(list(indices))
        self._next = batch_source(x[idx], y[idx], batch_size, seed=seed)
        self.device = torch.device(device)
        torch.manual_seed(0)
        self.model = SmallCNN().to(self.device)
        self.optimizer = torch.optim.SGD(self.model.parameters(), lr=lr, momentum=momentum)
        self.criterion = nn.CrossEntropyLoss()

    def next_batch(self):
        x, y = self._next()
        return x.to(self.device), y.to(self.device)

    def local_honest_gradient(self, *, x, y):
        self.model.zero_grad(set_to_none=True)
        self.criterion(self.model(x.to(self.device)), y.to(self.device)).backward()
        return flatten_grads(self.model)

    def apply_server_gradient(self, aggregated_grad):
        write_vector_to_grads_(self.model, aggregated_grad.to(self.device))
        self.optimizer.step()

    def dump_state_dict(self):
        return {k: v.detach().cpu() for k, v in self.model.state_dict().items()}


class DistributedPSByzNode(DistributedByzantineNode):
    def __init__(self, *, device: str = "cpu", scale: float = -1.0, pool_backend: str = "thread"):
        super().__init__(actor_pool=[ActorPoolConfig(backend=pool_backend, count=1, name="worker")],
                         attack=EmpireAttack(scale=scale), name=f"byz-{pool_backend}")
        self.device = torch.device(device)

    def next_batch(self):
        return torch.empty(0), torch.empty(0, dtype=torch.long)

    def apply_server_gradient(self, aggregated_grad):
        pass


__all__ = ["DistributedPSHonestNode", "DistributedPSByzNode", "select_pool_backend"]
