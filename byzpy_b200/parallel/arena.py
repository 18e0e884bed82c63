"""Flat parameter / gradient arenas.

The reference flattens gradients with ``torch.cat([p.grad.view(-1) ...])`` on
every step and writes the aggregate back with one ``copy_`` per parameter
(reference engine/node/mixin.py:15-24, examples/ps/nodes.py:17-33).  Here every
``nn.Parameter`` (and its ``.grad``) is re-homed once as a *view* into a flat
fp32 buffer, in ``module.parameters()`` order, tightly packed (same element
order as the reference's flat vector; only the tail is zero-padded), so
flatten/unflatten are no-ops and the fused kernels address whole
replicas through a single base pointer.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

PAD = 1024  # arenas are padded to a multiple of this many elements (vector width x ranks)


def flat_size(module: nn.Module) -> int:
    """Total number of parameters of ``module``."""
    return sum(p.numel() for p in module.parameters())


def padded_size(d: int, pad: int = PAD) -> int:
    """``d`` rounded up to a multiple of ``pad`` (arenas are padded so that every rank's shard is vector aligned)."""
    return (d + pad - 1) // pad * pad


class ParamArena:
    """Binds ``module``'s parameters and gradients to flat buffers.

    ``flat_params`` / ``flat_grads`` are 1-D fp32 tensors of ``d_pad`` elements;
    elements ``[d, d_pad)`` are padding and stay zero.
    """

    def __init__(self, module: nn.Module, *, flat_params: Optional[torch.Tensor] = None,
                 flat_grads: Optional[torch.Tensor] = None, pad: int = PAD):
        params = [p for p in module.parameters()]
        if not params:
            raise ValueError("module has no parameters")
        dev = params[0].device
        self.module = module
        self.d = sum(p.numel() for p in params)
        self.d_pad = padded_size(self.d, pad)
        self.offsets: List[Tuple[int, int, torch.Size]] = []
        if flat_params is None:
            flat_params = torch.zeros(self.d_pad, dtype=torch.float32, device=dev)
        if flat_grads is None:
            flat_grads = torch.zeros(self.d_pad, dtype=torch.float32, device=dev)
        if flat_params.numel() < self.d_pad or flat_grads.numel() < self.d_pad:
            raise ValueError("flat buffers are smaller than the padded parameter count")
        self.flat_params = flat_params[: self.d_pad]
        self.flat_grads = flat_grads[: self.d_pad]
        off = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                if p.dtype != torch.float32:
                    raise TypeError("ParamArena expects fp32 master parameters")
                dst = self.flat_params[off: off + n].view(p.shape)
                dst.copy_(p.detach())
                p.data = dst
                g = self.flat_grads[off: off + n].view(p.shape)
                g.zero_()
                p.grad = g
                self.offsets.append((off, n, p.shape))
                off += n
            if self.d_pad > self.d:
                self.flat_params[self.d:].zero_()
                self.flat_grads[self.d:].zero_()

    def zero_grad(self) -> None:
        self.flat_grads.zero_()

    def grad_vector(self) -> torch.Tensor:
        """The reference-layout flat gradient (a view, no copy)."""
        return self.flat_grads[: self.d]

    def param_vector(self) -> torch.Tensor:
        return self.flat_params[: self.d]

    def check_bound(self) -> bool:
        """True while every parameter / gradient still aliases the arena."""
        base_p = self.flat_params.data_ptr()
        base_g = self.flat_grads.data_ptr()
        for p, (off, n, _) in zip(self.module.parameters(), self.offsets):
            if p.data_ptr() != base_p + 4 * off:
                return False
            if p.grad is None or p.grad.data_ptr() != base_g + 4 * off:
                return False
        return True


def flatten_grads(module: nn.Module) -> torch.Tensor:
    """Reference-compatible flatten (zeros for missing grads); copies."""
    parts = []
    for p in module.parameters():
        parts.append((torch.zeros_like(p) if p.grad is None else p.grad).reshape(-1))
    return torch.cat(parts)


def flatten_params(module: nn.Module) -> torch.Tensor:
    """The parameters of ``module`` as one flat vector (a copy, ``module.parameters()`` order)."""
    return torch.cat([p.detach().reshape(-1) for p in module.parameters()])


def write_vector_to_params_(module: nn.Module, vec: torch.Tensor) -> None:
    """Load a flat vector into the parameters of ``module`` in place."""
    off = 0
    with torch.no_grad():
        for p in module.parameters():
            n = p.numel()
            p.copy_(vec[off: off + n].view_as(p))
            off += n


def write_vector_to_grads_(module: nn.Module, vec: torch.Tensor) -> None:
    """Load a flat vector into the ``.grad`` buffers of ``module`` (allocating them when absent), ready for ``optimizer.step()``."""
    off = 0
    for p in module.parameters():
        n = p.numel()
        chunk = vec[off: off + n].view_as(p).to(p.device)
        if p.grad is None:
            p.grad = chunk.clone()
        else:
            p.grad.copy_(chunk)
        off += n


__all__ = ["ParamArena", "flat_size", "padded_size", "flatten_grads", "flatten_params",
           "write_vector_to_params_", "write_vector_to_grads_", "PAD"]
