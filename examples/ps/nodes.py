"""Canonical user node classes for the parameter-server examples (counterpart of the reference's
examples/ps/nodes.py): an honest SmallCNN worker and an Empire-attack Byzantine worker built on the
``Distributed*Node`` bases, so aggregation / gradient / attack run through per-node pipelines."""
from __future__ import annotations

from typing import Sequence, Type

import torch
import torch.nn as nn

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian
from byzpy_b200.attacks import EmpireAttack
from byzpy_b200.engine.graph.pool import ActorPoolConfig
from byzpy_b200.engine.node.distributed import DistributedByzantineNode, DistributedHonestNode
from byzpy_b200.models import SmallCNN
from byzpy_b200 import ops
from byzpy_b200.parallel.arena import ParamArena, flatten_grads, write_vector_to_grads_  # noqa: F401  (re-exported)
from byzpy_b200.utils.data import batch_source, mnist_like


def select_pool_backend(spec: str) -> str:
    if spec == "process" or spec.startswith("tcp://"):
        return "thread"
    if spec.startswith("ucx://"):
        return "gpu"
    return spec


class DistributedPSHonestNode(DistributedHonestNode):
    def __init__(self, *, indices: Sequence[int], batch_size: int = 64, shuffle: bool = True, lr: float = 0.05,
                 momentum: float = 0.9, device: str = "cpu", data_root: str = "./data",
                 pool_backend: str = "thread", model_cls: Type[nn.Module] = SmallCNN, seed: int = 0):
        super().__init__(actor_pool=[ActorPoolConfig(backend=pool_backend, count=1, name="worker")],
                         aggregator=CoordinateWiseMedian(), metadata={"pool_backend": pool_backend},
                         name=f"honest-{pool_backend}")
        x, y = mnist_like(6000, root=data_root)
        idx = torch.as_tensor(list(indices))
        self._next = batch_source(x[idx], y[idx], batch_size, seed=seed, shuffle=shuffle)
        self.device = torch.device(device)
        torch.manual_seed(0)
        self.model = model_cls().to(self.device)
        # parameters and gradients live as views of two flat buffers: the flat gradient the server wants is
        # the buffer itself (no per-parameter cat), the aggregate is applied by one flat SGD(+momentum) update
        self.arena = ParamArena(self.model)
        self.lr, self.momentum = float(lr), float(momentum)
        self.velocity = torch.zeros_like(self.arena.flat_params) if momentum else None
        self.criterion = nn.CrossEntropyLoss()

    def next_batch(self):
        x, y = self._next()
        return x.to(self.device), y.to(self.device)

    def local_honest_gradient(self, *, x, y):
        self.arena.zero_grad()
        self.criterion(self.model(x.to(self.device)), y.to(self.device)).backward()
        return self.arena.grad_vector().clone()

    def apply_server_gradient(self, aggregated_grad):
        d = self.arena.d
        g = aggregated_grad.reshape(-1).to(device=self.device, dtype=torch.float32)
        ops.sgd_step(g, [self.arena.flat_params[:d]], [self.velocity[:d]] if self.velocity is not None else None,
                     lr=self.lr, momentum=self.momentum)

    def dump_state_dict(self):
        return {k: v.detach().cpu() for k, v in self.model.state_dict().items()}


class DistributedPSByzNode(DistributedByzantineNode):
    def __init__(self, *, device: str = "cpu", scale: float = -1.0, pool_backend: str = "thread"):
        super().__init__(actor_pool=[ActorPoolConfig(backend=pool_backend, count=1, name="worker")],
                         attack=EmpireAttack(scale=scale), metadata={"pool_backend": pool_backend},
                         name=f"byz-{pool_backend}")
        self.device = torch.device(device)

    def next_batch(self):
        return torch.empty(0), torch.empty(0, dtype=torch.long)

    def apply_server_gradient(self, aggregated_grad):
        pass


__all__ = ["DistributedPSHonestNode", "DistributedPSByzNode", "select_pool_backend"]
