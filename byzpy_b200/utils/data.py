"""Small data helpers for the examples: MNIST when a local copy exists (no download -- the
build/CI boxes have no network), otherwise a deterministic synthetic stand-in of the same shape
whose labels are a fixed random linear function of the pixels (so models can actually learn).
Counterpart of the data plumbing inside reference examples/ps/thread/mnist.py:30-54 (strided shards,
``evaluate``)."""
from __future__ import annotations

import os
from typing import Callable, List, Tuple

import torch


def mnist_like(n: int = 6000, *, train: bool = True, root: str = "./data", seed: int = 0):
    """Returns ``(images float32 (n,1,28,28) in [0,1], labels int64 (n,))``."""
    try:
        from torchvision import datasets

        if os.path.isdir(os.path.join(root, "MNIST", "raw")):
            ds = datasets.MNIST(root=root, train=train, download=False)
            x = ds.data[:n].float().div_(255.0).unsqueeze(1)
            return x, ds.targets[:n].clone()
    except Exception:
        pass
    g = torch.Generator().manual_seed(seed if train else seed + 1)
    protos = torch.rand(10, 1, 28, 28, generator=torch.Generator().manual_seed(1234))
    y = torch.randint(0, 10, (n,), generator=g)
    x = (protos[y] + 0.35 * torch.randn(n, 1, 28, 28, generator=g)).clamp_(0, 1)
    return x, y


def shard_indices(n_items: int, n_shards: int) -> List[List[int]]:
    """Round-robin split of ``range(n_items)`` into ``n_shards`` index lists (shard ``i`` gets ``i, i + n_shards, ...``)."""
    return [list(range(i, n_items, n_shards)) for i in range(n_shards)]


def batch_source(x: torch.Tensor, y: torch.Tensor, batch_size: int, *, seed: int = 0,
                 pin: bool = False, shuffle: bool = True) -> Callable[[], Tuple[torch.Tensor, torch.Tensor]]:
    """Infinite mini-batch source over in-memory tensors (a DataLoader without workers): reshuffled every
    epoch, or cycling in storage order with ``shuffle=False``.  Incomplete trailing batches are dropped."""
    g = torch.Generator().manual_seed(seed)

    def order():
        return torch.randperm(x.shape[0], generator=g) if shuffle else torch.arange(x.shape[0])

    state = {"perm": order(), "pos": 0}

    def next_batch():
        if state["pos"] + batch_size > x.shape[0]:
            state["perm"] = order()
            state["pos"] = 0
        idx = state["perm"][state["pos"]: state["pos"] + batch_size]
        state["pos"] += batch_size
        xb, yb = x[idx], y[idx]
        return (xb.pin_memory(), yb.pin_memory()) if pin and torch.cuda.is_available() else (xb, yb)

    return next_batch


@torch.no_grad()
def evaluate(model: torch.nn.Module, x: torch.Tensor, y: torch.Tensor, device, batch: int = 1024):
    """``(mean loss, accuracy)`` of ``model`` on the tensors ``x, y``, evaluated in batches of ``batch`` on ``device``."""
    model.eval()
    loss_sum, correct = 0.0, 0
    for i in range(0, x.shape[0], batch):
        xb, yb = x[i:i + batch].to(device), y[i:i + batch].to(device)
        out = model(xb)
        loss_sum += torch.nn.functional.cross_entropy(out, yb, reduction="sum").item()
        correct += (out.argmax(1) == yb).sum().item()
    model.train()
    return loss_sum / x.shape[0], correct / x.shape[0]


__all__ = ["mnist_like", "shard_indices", "batch_source", "evaluate"]
