"""P2P example nodes (counterpart of the reference's examples/p2p/nodes.py): SmallCNN honest node
with the P2P mixin, Empire Byzantine node.

``P2PHonestNode`` / ``P2PByzNode`` are the light versions (mixin only, no per-node actor pool);
``DistributedP2PHonestNode`` / ``DistributedP2PByzNode`` carry the reference's names and constructor
arguments and additionally are ``Distributed*Node`` s, i.e. own a pool and the aggregate / attack pipelines."""
from __future__ import annotations

from typing import Sequence, Type

import torch
import torch.nn as nn

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseTrimmedMean
from byzpy_b200.attacks import EmpireAttack
from byzpy_b200.engine.graph.pool import ActorPoolConfig
from byzpy_b200.engine.node.distributed import DistributedByzantineNode, DistributedHonestNode
from byzpy_b200.engine.node.mixin import P2PByzantineMixin, P2PHonestMixin
from byzpy_b200.models import SmallCNN
from byzpy_b200.parallel.arena import ParamArena
from byzpy_b200.utils.data import batch_source, mnist_like


class P2PHonestNode(P2PHonestMixin):
    def __init__(self, *, indices: Sequence[int], batch_size: int = 64, device: str = "cpu", f: int = 1, seed: int = 0,
                 data_root: str = "./data"):
        x, y = mnist_like(6000, root=data_root)
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


class DistributedP2PHonestNode(P2PHonestMixin, DistributedHonestNode):
    def __init__(self, *, indices: Sequence[int], batch_size: int = 64, shuffle: bool = True, lr: float = 0.05,
                 momentum: float = 0.9, device: str = "cpu", data_root: str = "./data",
                 pool_backend: str = "thread", model_cls: Type[nn.Module] = SmallCNN, f: int = 1, seed: int = 0):
        agg = CoordinateWiseTrimmedMean(f=f)
        DistributedHonestNode.__init__(self, actor_pool=[ActorPoolConfig(backend=pool_backend, count=1, name="worker")],
                                       aggregator=agg, metadata={"pool_backend": pool_backend},
                                       name=f"p2p-honest-{pool_backend}")
        x, y = mnist_like(6000, root=data_root)
        idx = torch.as_tensor(list(indices))
        self._next = batch_source(x[idx], y[idx], batch_size, seed=seed, shuffle=shuffle)
        self.device = torch.device(device)
        torch.manual_seed(0)
        self.model = model_cls().to(self.device)
        self.arena = ParamArena(self.model)
        self.criterion = nn.CrossEntropyLoss()
        self.optimizer = torch.optim.SGD(self.model.parameters(), lr=lr, momentum=momentum)
        self.lr = float(lr)
        self.p2p_agg, self.p2p_pre = agg, None

    def next_batch(self):
        x, y = self._next()
        return x.to(self.device), y.to(self.device)

    def local_honest_gradient(self, *, x, y):
        self.arena.zero_grad()
        self.criterion(self.model(x), y).backward()
        return self.arena.grad_vector().clone()

    def apply_server_gradient(self, aggregated_grad):
        with torch.no_grad():
            self.arena.param_vector().sub_(aggregated_grad.reshape(-1).to(self.device), alpha=self.lr)

    def dump_state_dict(self):
        return {k: v.detach().cpu() for k, v in self.model.state_dict().items()}


class DistributedP2PByzNode(P2PByzantineMixin, DistributedByzantineNode):
    def __init__(self, *, device: str = "cpu", scale: float = -1.0, pool_backend: str = "thread"):
        self.device = torch.device(device)
        DistributedByzantineNode.__init__(self, actor_pool=[ActorPoolConfig(backend=pool_backend, count=1, name="worker")],
                                          attack=EmpireAttack(scale=scale), metadata={"pool_backend": pool_backend},
                                          name=f"p2p-byz-{pool_backend}")

    def next_batch(self):
        return torch.empty(0), torch.empty(0, dtype=torch.long)

    def apply_server_gradient(self, aggregated_grad):
        return None


__all__ = ["P2PHonestNode", "P2PByzNode", "DistributedP2PHonestNode", "DistributedP2PByzNode"]
