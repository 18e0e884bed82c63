"""Gaussian attack: an i.i.d. N(mu, sigma^2) vector shaped like a gradient, re-seeded every call
so a fixed ``seed`` reproduces the same vector (reference attacks/gaussian.py:38-139).

CPU inputs use ``numpy.random.default_rng(seed)`` exactly like the reference (bit-identical);
CUDA inputs are sampled on the device with a Philox4x32-10 kernel (statistically equivalent,
SURVEY K18)."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import ops
from ..aggregators._chunking import select_adaptive_chunk_size
from ..aggregators.base import finish, pool_size_of, prepare_rows
from ..engine.graph.subtask import SubTask
from .base import Attack


def _span(start: int, end: int):
    return start, end


class GaussianAttack(Attack):
    """Gaussian noise: submit a vector of i.i.d. ``N(mu, sigma^2)`` samples shaped like a gradient.

    Parameters
    ----------
    mu, sigma : float, default 0.0, 1.0
        Mean and standard deviation of every coordinate (``sigma >= 0``).
    seed : int, optional
        The generator is re-seeded with it on every call, so a fixed seed submits the same vector each round; ``None``
        draws fresh noise.
    chunk_size : int, default 8192
        Coordinates per subtask on an actor pool.

    Notes
    -----
    ``honest_grads`` is only used for shape, dtype and device.  CPU inputs are sampled with
    ``numpy.random.default_rng(seed)`` (bit-identical with the reference); CUDA inputs with a counter-based
    Philox4x32-10 kernel on the device, which has the same distribution but not the same bits.

    Examples
    --------
    >>> import torch
    >>> from byzpy_b200.attacks import GaussianAttack
    >>> a = GaussianAttack(mu=0.0, sigma=1.0, seed=7).apply(honest_grads=[torch.zeros(1000)])
    >>> b = GaussianAttack(mu=0.0, sigma=1.0, seed=7).apply(honest_grads=[torch.zeros(1000)])
    >>> a.shape, bool(torch.equal(a, b)), bool(abs(a.std().item() - 1.0) < 0.1)
    (torch.Size([1000]), True, True)
    """

    name = "gaussian"
    max_subtasks_inflight = 0       # 0 / None: the pool-sized default window (value of the reference class)
    uses_honest_grads = True
    supports_subtasks = True

    def __init__(self, mu: float = 0.0, sigma: float = 1.0, *, seed: Optional[int] = None,
                 chunk_size: int = 8192) -> None:
        if sigma < 0:
            raise ValueError("sigma must be >= 0")
        if chunk_size <= 0:
            raise ValueError("chunk_size must be > 0")
        self.mu, self.sigma, self.seed = float(mu), float(sigma), seed
        self.chunk_size = int(chunk_size)

    def apply(self, *, model=None, x=None, y=None, honest_grads=None, base_grad=None):
        if not honest_grads:
            raise ValueError("GaussianAttack requires honest_grads.")
        rows, like = prepare_rows([honest_grads[0]], "honest_grads")
        d = rows[0].numel()
        if like.is_cuda:
            seed = self.seed if self.seed is not None else int(np.random.SeedSequence().entropy % (2 ** 63))
            out = torch.empty(d, dtype=torch.float32, device=like.device)
            ops.gaussian_(out, self.mu, self.sigma, seed)
            return finish(out, like)
        sample = np.random.default_rng(self.seed).normal(loc=self.mu, scale=self.sigma, size=d)
        return finish(torch.from_numpy(sample), like)

    # -- subtask path --------------------------------------------------------------------------------
    # A seeded PCG64 normal stream cannot be entered in the middle (the ziggurat sampler draws a variable
    # number of words per sample), and "same seed -> same vector" is the contract, so the chunk subtasks
    # only lay out the output spans (as the reference's no-op chunks do, gaussian.py:33-35, 100-137) and
    # the vector is drawn once in the reduce step.
    def create_subtasks(self, inputs, *, context):
        grads = inputs.get("honest_grads")
        if not grads:
            raise ValueError("GaussianAttack requires honest_grads.")
        rows, _ = prepare_rows([grads[0]], "honest_grads")
        d = rows[0].numel()
        chunk = select_adaptive_chunk_size(d, self.chunk_size, pool_size=pool_size_of(context))
        return [SubTask(fn=_span, args=(s, min(d, s + chunk)), name=f"gaussian_chunk_{k}")
                for k, s in enumerate(range(0, d, chunk))]

    def reduce_subtasks(self, partials, inputs, *, context):
        if partials:
            d = prepare_rows([inputs["honest_grads"][0]], "honest_grads")[0][0].numel()
            spans = sorted(partials)
            if spans[0][0] != 0 or spans[-1][1] != d or any(a[1] != b[0] for a, b in zip(spans, spans[1:])):
                raise ValueError("gaussian chunk spans do not tile the gradient")
        return self.compute(inputs, context=context)


__all__ = ["GaussianAttack"]
