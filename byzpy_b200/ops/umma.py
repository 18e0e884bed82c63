"""Front-end of the tcgen05 / TMEM Gram kernel (``csrc/gram_umma.cu``) -- the ``flat @ flat.T`` of
reference krum.py:41-44 / nnm.py:87-88 on the 5th-generation tensor cores.

The tensor-core kernel consumes whole tiles of ``32 * (128 / n_pad)`` columns (block-diagonal
packing, 16-byte aligned 128-byte row segments); the tail (< one tile) goes through the exact fp32
CUDA-core kernel and is added in the fp64 reduction.  Rows that are not 16-byte aligned fall
back to the CUDA-core kernel entirely.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _gram_scratch, _stream, require_ext, sm_count

_SCRATCH: dict = {}


def _partials(dev: torch.device, n: int, slots: int) -> torch.Tensor:
    need = max(1, slots) * 2 * n * n
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.float32, device=dev)
        _SCRATCH[key] = buf
    return buf


def supported(rows: Sequence[torch.Tensor]) -> bool:
    return all(r.data_ptr() % 16 == 0 for r in rows)


def gram_umma(rows: List[torch.Tensor], scales: List[float], G: torch.Tensor,
              G64: Optional[torch.Tensor]) -> None:
    ext = require_ext()
    n = len(rows)
    d = rows[0].numel()
    dev = rows[0].device
    ptrs = [r.data_ptr() for r in rows]
    sms = sm_count(dev)
    stream = _stream(dev)
    tc = ext.gram_umma_tile_cols(n)
    if not supported(rows) or d < tc:
        scratch = _gram_scratch(dev, n)
        ext.gram(ptrs, scales, 0, d, scratch.data_ptr(), scratch.numel() // (n * n), G.data_ptr(),
                 G64.data_ptr() if G64 is not None else 0, sms, stream)
        return
    main = (d // tc) * tc
    tail64 = None
    if main < d:
        scratch = _gram_scratch(dev, n)
        tail32 = torch.empty((n, n), dtype=torch.float32, device=dev)
        tail64 = torch.empty((n, n), dtype=torch.float64, device=dev)
        ext.gram(ptrs, scales, main, d - main, scratch.data_ptr(), scratch.numel() // (n * n),
                 tail32.data_ptr(), tail64.data_ptr(), sms, stream)
    grid = ext.gram_umma_grid(n, main, sms)
    part = _partials(dev, n, ext.gram_umma_partials(n, grid))
    ext.gram_umma(ptrs, scales, 0, main, part.data_ptr(), part.numel() // (2 * n * n),
                  tail64.data_ptr() if tail64 is not None else 0, G.data_ptr(),
                  G64.data_ptr() if G64 is not None else 0, sms, stream)


__all__ = ["gram_umma", "supported"]
