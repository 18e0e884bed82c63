"""Front-end of the tcgen05 / TMEM Gram kernel (``csrc/gram_umma.cu``) -- the ``flat @ flat.T`` of
reference krum.py:41-44 / nnm.py:87-88 on the 5th-generation tensor cores.

The tensor-core kernel consumes whole tiles of ``32 * (128 / n_pad)`` columns (block-diagonal
packing, 16-byte aligned 128-byte row segments); the tail (< one tile) goes through the exact fp32
CUDA-core kernel and is added in the fp64 reduction.  Rows that are not 16-byte aligned fall
back to the CUDA-core kernel entirely.
"""
from __future__ import annotations

import os
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
    """The tcgen05 Gram kernel stages rows with 16-byte copies: every row must start on a 16-byte boundary."""
    return all(r.data_ptr() % 16 == 0 for r in rows)


# ---------------------------------------------------------------------------- TMA segments
MAX_SEGMENTS = 12
_MAPS: dict = {}
_TMA_OK: Optional[bool] = None


def tma_available() -> bool:
    """The TMA-fed kernel needs the driver's tensor-map encoder (dlopen'ed libcuda)."""
    global _TMA_OK
    if _TMA_OK is None:
        ext = require_ext()
        _TMA_OK = bool(hasattr(ext, "gram_umma_tma") and ext.vmm_support(0)["reason"] == ""
                       and os.environ.get("BYZPY_GRAM_TMA", "1") != "0")
    return _TMA_OK


def segments_of(ptrs: Sequence[int]) -> Optional[List[tuple]]:
    """Group a row pointer table into row-major matrix segments ``(base, rows, stride_bytes)``:
    maximal runs of rows at a constant positive byte stride (multiple of 16).  A stacked ``(n, d)``
    tensor or a flat arena is one segment; the fused round's table is one segment per peer GPU.
    None when the table does not fit :data:`MAX_SEGMENTS` segments."""
    segs: List[tuple] = []
    i, n = 0, len(ptrs)
    while i < n:
        base = ptrs[i]
        if base % 16:
            return None
        rows, stride = 1, 0
        if i + 1 < n:
            st = ptrs[i + 1] - base
            if st > 0 and st % 16 == 0:
                stride = st
                rows = 2
                while i + rows < n and ptrs[i + rows] - ptrs[i + rows - 1] == st and rows < 256:
                    rows += 1
        segs.append((base, rows, stride))
        i += rows
        if len(segs) > MAX_SEGMENTS:
            return None
    return segs


def tma_maps(ptrs: Sequence[int], row_len: int) -> Optional[bytes]:
    """Tensor maps (opaque blob, cached) for the TMA-fed Gram kernel over these rows, or None when
    the rows do not form few enough segments / the encoder is unavailable.  ``row_len``: elements
    addressable from every row pointer (the map's inner dimension)."""
    if not tma_available():
        return None
    segs = segments_of(ptrs)
    if segs is None:
        return None
    for _, rows, stride in segs:
        if rows > 1 and stride < row_len * 4:
            return None                     # overlapping rows: not a matrix
    key = (tuple(segs), int(row_len), len(ptrs))
    blob = _MAPS.get(key)
    if blob is None:
        if len(_MAPS) > 512:
            _MAPS.clear()
        try:
            blob = require_ext().gram_tma_maps(segs, int(row_len), len(ptrs))
        except Exception:
            blob = False
        _MAPS[key] = blob
    return blob or None


def launch(ext, ptrs: Sequence[int], scales, off: int, main: int, row_len: int, partials: torch.Tensor, n: int,
           tail_ptr: int, G_ptr: int, G64_ptr: int, sms: int, stream: int) -> str:
    """Run the tcgen05 Gram over columns ``[off, off + main)``: TMA-fed when the rows form matrix
    segments, per-thread cp.async from the pointer table otherwise.  Returns which path ran."""
    slots = partials.numel() // (2 * n * n)
    maps = tma_maps(ptrs, row_len)
    if maps is not None:
        ext.gram_umma_tma(maps, list(ptrs), scales, off, main, partials.data_ptr(), slots, tail_ptr, G_ptr, G64_ptr,
                          sms, stream)
        return "tma"
    ext.gram_umma(list(ptrs), scales, off, main, partials.data_ptr(), slots, tail_ptr, G_ptr, G64_ptr, sms, stream)
    return "cp.async"


def gram_umma(rows: List[torch.Tensor], scales: List[float], G: torch.Tensor,
              G64: Optional[torch.Tensor]) -> None:
    """Gram matrix of ``rows`` (each scaled by ``scales[i]``) on the 5th-generation tensor cores.

    fp32 operands are split into three TF32 terms (3xTF32) so the products carry fp32 accuracy; accumulators live in
    TMEM, tiles arrive by TMA (``cp.async`` staging when the rows are not one strided matrix), split-K partials are
    reduced in fp64.  Writes the fp32 result into ``G`` and, when given, the fp64 one into ``G64``.
    """
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
    global last_path
    last_path = launch(ext, ptrs, scales, 0, main, d, part, n, tail64.data_ptr() if tail64 is not None else 0,
                       G.data_ptr(), G64.data_ptr() if G64 is not None else 0, sms, stream)


last_path = ""       # "tma" | "cp.async": which loader the last gram_umma() call used (tests, benchmarks)

__all__ = ["gram_umma", "supported", "segments_of", "tma_maps", "tma_available", "launch"]
