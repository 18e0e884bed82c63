"""Device-resident nodes: the ``"gpu"`` actor backend of this framework is a CUDA runtime.

A ``DeviceHonestNode`` / ``DeviceByzantineNode`` satisfies the ordinary node contract
(``next_batch`` / ``honest_gradient`` / ``apply_server_gradient``), so it works with the generic
actor-based :class:`ParameterServer` path, and additionally exposes a :class:`DeviceWorker` so
that a ``ParameterServer`` made only of device nodes runs its round as ONE fused kernel over
flat arenas (see :mod:`byzpy_b200.parallel.device_ps`).

Reference analogue: the example node classes (reference examples/ps/nodes.py:64-161) hosted on
``GPUActorBackend`` (reference engine/actor/backends/gpu.py:23-200), which contains no CUDA code.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.nn as nn

from ...attacks.base import Attack
from ...parallel.arena import flatten_grads, write_vector_to_grads_
from ...parallel.device_ps import DeviceWorker, RowFold
from .base import ByzantineNode, HonestNode
from .mixin import P2PByzantineMixin, P2PHonestMixin

BatchSource = Callable[[], Tuple[torch.Tensor, torch.Tensor]]


class DeviceHonestNode(HonestNode):
    """An honest worker that lives on a GPU: model, loss, optimizer state and data source.

    With nodes of this kind :class:`~byzpy_b200.engine.parameter_server.ps.ParameterServer` takes its fused device path:
    all replicas of a GPU run forward and backward in one captured CUDA graph and write their gradients into rows of a
    symmetric-memory arena, where the fused kernels aggregate them.  The same object also satisfies the generic node
    contract (``next_batch`` / ``honest_gradient`` / ``apply_server_gradient``), so it works through actors on any
    backend, CPU included.

    Parameters
    ----------
    model : torch.nn.Module
    loss_fn : callable, optional
        Default cross entropy.
    data : callable, optional
        ``() -> (inputs, targets)``; may return host tensors (pinned memory makes the copy asynchronous).
    lr, momentum, weight_decay : float
        SGD hyper-parameters applied by the fused optimizer step (or by ``ensure_optimizer()`` on the generic path).
    device : str, optional
        Default ``"cuda"`` when available.
    preprocess : callable, optional
        Applied to the inputs on the device (normalisation, layout change) inside the captured graph.
    name : str
    """

    def __init__(self, model: nn.Module, loss_fn: Optional[Callable] = None, *,
                 data: Optional[BatchSource] = None, lr: float = 0.05, momentum: float = 0.9,
                 weight_decay: float = 0.0, device: Optional[str] = None,
                 preprocess: Optional[Callable] = None, name: str = "honest"):
        self.device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        self.model = model.to(self.device)
        self.loss_fn = loss_fn or nn.CrossEntropyLoss()
        self.data = data
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.preprocess = preprocess
        self.name = name
        self._opt: Optional[torch.optim.Optimizer] = None
        self.worker = DeviceWorker(self.model, self.loss_fn, role="honest", name=name,
                                   preprocess=preprocess, data=data)

    # ---- generic node contract (used by the actor-based round) ----------------------------
    def next_batch(self):
        if self.data is None:
            raise RuntimeError("DeviceHonestNode has no data source")
        x, y = self.data()
        return x.to(self.device, non_blocking=True), y.to(self.device, non_blocking=True)

    def honest_gradient(self, x, y):
        self.model.zero_grad(set_to_none=False)
        xin = self.preprocess(x) if self.preprocess is not None else x
        loss = self.loss_fn(self.model(xin), y)
        loss.backward()
        return flatten_grads(self.model)

    def ensure_optimizer(self) -> torch.optim.Optimizer:
        """The node's SGD optimizer (created on first use; checkpoint loading needs it to exist)."""
        if self._opt is None:
            self._opt = torch.optim.SGD(self.model.parameters(), lr=self.lr, momentum=self.momentum,
                                        weight_decay=self.weight_decay)
        return self._opt

    def apply_server_gradient(self, grad_vec):
        opt = self.ensure_optimizer()
        write_vector_to_grads_(self.model, grad_vec.to(self.device))
        opt.step()

    def dump_state_dict(self):
        return {k: v.detach().cpu() for k, v in self.model.state_dict().items()}


class DeviceByzantineNode(ByzantineNode):
    """Byzantine node driven by an :class:`Attack`.

    Attacks that need the node's own gradient (``uses_base_grad``, e.g. SignFlip) require a
    ``model`` + ``data``; omniscient attacks (Little, Empire, Mimic) need neither and become
    virtual/alias rows of the fused kernel.
    """

    def __init__(self, attack: Attack, *, model: Optional[nn.Module] = None,
                 loss_fn: Optional[Callable] = None, data: Optional[BatchSource] = None,
                 lr: float = 0.05, momentum: float = 0.9, weight_decay: float = 0.0,
                 device: Optional[str] = None, preprocess: Optional[Callable] = None,
                 name: str = "byzantine"):
        self.attack = attack
        self.device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        self.model = model.to(self.device) if model is not None else None
        self.loss_fn = loss_fn or nn.CrossEntropyLoss()
        self.data = data
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.preprocess = preprocess
        self.name = name
        self._opt = None
        self.worker: Optional[DeviceWorker] = None
        if attack.uses_base_grad or attack.uses_model_batch:
            if model is None:
                raise ValueError(f"{type(attack).__name__} needs the node's own model/batch")
            self.worker = DeviceWorker(self.model, self.loss_fn, role="byzantine", name=name,
                                       preprocess=preprocess, data=data)

    def fold(self, n_honest: int) -> Optional[RowFold]:
        return self.attack.fold(n_honest)

    def next_batch(self):
        if self.data is None:
            return torch.empty(0), torch.empty(0, dtype=torch.long)
        x, y = self.data()
        return x.to(self.device, non_blocking=True), y.to(self.device, non_blocking=True)

    def _own_gradient(self, x, y):
        self.model.zero_grad(set_to_none=False)
        xin = self.preprocess(x) if self.preprocess is not None else x
        self.loss_fn(self.model(xin), y).backward()
        return flatten_grads(self.model)

    def byzantine_gradient(self, x, y, honest_grads=None):
        kw = {}
        if self.attack.uses_model_batch:
            if x.numel() == 0:
                x, y = self.next_batch()
            kw.update(model=self.model, x=self.preprocess(x) if self.preprocess else x, y=y)
        if self.attack.uses_base_grad:
            if x.numel() == 0:
                x, y = self.next_batch()
            kw["base_grad"] = self._own_gradient(x, y)
        if self.attack.uses_honest_grads:
            kw["honest_grads"] = list(honest_grads or [])
        return self.attack.apply(**kw)

    def ensure_optimizer(self) -> Optional[torch.optim.Optimizer]:
        if self.model is not None and self._opt is None:
            self._opt = torch.optim.SGD(self.model.parameters(), lr=self.lr, momentum=self.momentum,
                                        weight_decay=self.weight_decay)
        return self._opt

    def apply_server_gradient(self, grad_vec):
        if self.model is None:
            return
        opt = self.ensure_optimizer()
        write_vector_to_grads_(self.model, grad_vec.to(self.device))
        opt.step()


class DeviceP2PHonestNode(P2PHonestMixin):
    """Honest gossip node.  Generic path: the ``P2PHonestMixin`` step functions.  Device path:
    :class:`byzpy_b200.parallel.device_p2p.DeviceP2PRound` (fused, peer rows read over NVLink)."""

    def __init__(self, model: nn.Module, aggregator, *, loss_fn: Optional[Callable] = None,
                 pre_aggregator=None, data: Optional[BatchSource] = None, device: Optional[str] = None,
                 preprocess: Optional[Callable] = None, name: str = "p2p-honest"):
        self.device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        self.model = model.to(self.device)
        self.criterion = loss_fn or nn.CrossEntropyLoss()
        self.p2p_agg = aggregator
        self.p2p_pre = pre_aggregator
        self.data = data
        self.preprocess = preprocess
        self.name = name

    def next_batch(self):
        x, y = self.data()
        x, y = x.to(self.device, non_blocking=True), y.to(self.device, non_blocking=True)
        return (self.preprocess(x) if self.preprocess is not None else x), y

    def dump_state_dict(self):
        return {k: v.detach().cpu() for k, v in self.model.state_dict().items()}


class DeviceP2PByzantineNode(P2PByzantineMixin):
    """Byzantine gossip node for the device path: holds the :class:`~byzpy_b200.attacks.base.Attack` the fused P2P
    round folds into its kernels (or applies through ``p2p_broadcast_vector`` on the generic path).

    Parameters
    ----------
    attack : Attack
    device : str, optional
    name : str
    """

    def __init__(self, attack: Attack, *, device: Optional[str] = None, name: str = "p2p-byz"):
        self.device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        self.attack = attack
        self.name = name


__all__ = ["DeviceHonestNode", "DeviceByzantineNode", "DeviceP2PHonestNode", "DeviceP2PByzantineNode"]
