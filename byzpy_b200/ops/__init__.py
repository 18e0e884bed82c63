"""Functional front-end of the sm_100a kernel library.

Every function takes the logical ``(n, d)`` gradient matrix as a *list of n
row tensors* (or one 2-D tensor): rows are addressed through a pointer table,
so nothing is ever ``torch.stack``-ed on the GPU path -- rows may live in
different allocations, in a flat arena, or in a peer GPU's memory.

Dispatch rule (no silent fallbacks on a GPU box):
  * CUDA fp32 rows, ``n <= 128``  -> hand-written kernels in ``byzpy_b200._C``;
    if the extension cannot be imported a ``RuntimeError`` is raised.
  * CPU tensors (and exotic dtypes / n > 128) -> the plain PyTorch reference
    implementations in :mod:`byzpy_b200.ops.reference`, which are also the
    oracle the GPU numerics tests compare against.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple, Union

import torch

from . import reference as ref

MODE_MEDIAN, MODE_TRMEAN, MODE_MEAMED, MODE_MEAN = 0, 1, 2, 3
MAXN = 128
MAXR = 16

_C = None
_C_err: Optional[BaseException] = None


def _load_ext():
    global _C, _C_err
    if _C is not None:
        return _C
    try:
        from .. import _C as ext  # type: ignore[attr-defined]

        _C = ext
    except Exception as exc:  # pragma: no cover - exercised only on broken installs
        _C_err = exc
        _C = None
    return _C


# Kernel launches issued from Python-side layer wrappers (fused BN / pooling), so the training engine
# can report how many of this library's kernels a step contains (bench.py "gpu_launches").
_LAUNCHES = [0]


def count_launch(k: int = 1) -> None:
    """Record ``k`` launches of this library's kernels made from a Python-side layer wrapper."""
    _LAUNCHES[0] += k


def launches() -> int:
    """Kernel launches recorded so far by :func:`count_launch` (``bench.py`` reports the per-step difference)."""
    return _LAUNCHES[0]


def extension_available() -> bool:
    """True when the compiled sm_100a kernel library (``byzpy_b200._C``) can be imported."""
    return _load_ext() is not None


def require_ext():
    """The kernel library module, or ``RuntimeError`` when it is not built: CUDA inputs never fall back to PyTorch silently."""
    ext = _load_ext()
    if ext is None:
        raise RuntimeError(
            "byzpy_b200._C (sm_100a kernel library) is not built/importable; run "
            "`python -m byzpy_b200._build`. Refusing to fall back to PyTorch on a CUDA device. "
            f"Import error: {_C_err!r}"
        )
    return ext


_SM_COUNT: dict = {}


def sm_count(device: torch.device) -> int:
    """Number of SMs of ``device`` (cached); the persistent kernels size their grids from it."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


Rows = Union[torch.Tensor, Sequence[torch.Tensor]]


def as_rows(x: Rows) -> List[torch.Tensor]:
    """Normalise the input into a list of 1-D tensors (views, no copy)."""
    if isinstance(x, torch.Tensor):
        if x.dim() == 1:
            return [x]
        flat = x.reshape(x.shape[0], -1)
        return [flat[i] for i in range(flat.shape[0])]
    if type(x) is list and x and type(x[0]) is torch.Tensor:
        # already a list of flat tensors of one length (what the operator bases pass): one cheap scan
        d0 = x[0].numel()
        if all(type(r) is torch.Tensor and r.dim() == 1 and r.numel() == d0 for r in x):
            return x
    rows = []
    for r in x:
        if not isinstance(r, torch.Tensor):
            r = torch.as_tensor(r)
        rows.append(r if r.dim() == 1 else r.reshape(-1))
    if not rows:
        raise ValueError("need at least one row")
    d = rows[0].numel()
    for r in rows:
        if r.numel() != d:
            raise ValueError("all rows must have the same number of elements")
    return rows


def _kernel_ok(rows: List[torch.Tensor]) -> bool:
    r0 = rows[0]
    if not r0.is_cuda or len(rows) > MAXN:
        return False
    return all(r.is_cuda and r.dtype == torch.float32 and r.device == r0.device for r in rows)


def _prep(rows: List[torch.Tensor]) -> List[torch.Tensor]:
    return [r if r.is_contiguous() else r.contiguous() for r in rows]


def _scales(scales, n) -> List[float]:
    if scales is None:
        return []
    s = [float(v) for v in scales]
    if len(s) != n:
        raise ValueError("scales must have one entry per row")
    return s


# ----------------------------------------------------------------------------
# coordinate-wise family
# ----------------------------------------------------------------------------
def cw_select(
    rows: Rows,
    mode: int,
    f: int = 0,
    *,
    scales: Optional[Sequence[float]] = None,
    virtual: Optional[Tuple[int, int, float, float]] = None,
    out: Optional[torch.Tensor] = None,
    update: Optional[dict] = None,
    impl: str = "auto",
) -> torch.Tensor:
    """Coordinate-wise select over n rows.

    ``impl``: ``"auto"`` | ``"direct"`` (register loads) | ``"staged"`` (thread-private cp.async pipeline) |
    ``"tiled"`` (warp-tiled cp.async pipeline for 17..128 rows; shorter inputs and other shapes fall back to
    ``"auto"``).  The CUDA variants are bit-identical, the knob exists for benchmarking; ``"auto"`` can be
    redirected with ``BYZPY_CW_IMPL=tiled`` (the tiled kernel was written after the round's GPU time was spent:
    it is opt-in until a B200 run has timed it, see profiles/cw_select.md section 4).

    ``virtual=(count, n_honest, a, b)`` appends ``count`` synthesised rows equal
    to ``a*mean + b*std`` of the first ``n_honest`` rows (Little / Empire).
    ``update=dict(params=[...], moms=[...]|None, lr=, momentum=, weight_decay=)``
    fuses an SGD step on flat replicas into the same kernel.
    """
    rows = as_rows(rows)
    n = len(rows)
    d = rows[0].numel()
    nv, nh, va, vb = virtual if virtual is not None else (0, 0, 0.0, 0.0)
    nt = n + nv
    if mode == MODE_TRMEAN and not (0 <= 2 * f < nt):
        raise ValueError("f must satisfy 0 <= 2f < n")
    if mode == MODE_MEAMED and not (0 <= f < nt):
        raise ValueError("f must satisfy 0 <= f < n")
    if _kernel_ok(rows) and nt <= MAXN:
        ext = require_ext()
        rows = _prep(rows)
        dev = rows[0].device
        if out is None:
            out = torch.empty(d, dtype=torch.float32, device=dev)
        params, moms, lr, mu, wd = _unpack_update(update)
        ext.cw_select(
            [r.data_ptr() for r in rows], _scales(scales, n), mode, int(f), int(nv), int(nh),
            float(va), float(vb), 0, d, out.data_ptr(),
            [p.data_ptr() for p in params], [m.data_ptr() for m in moms],
            lr, mu, wd, sm_count(dev), _stream(dev), _CW_IMPL[_cw_impl(impl)],
        )
        return out
    res = _host_cw_select(rows, mode, f, scales, virtual, out)
    if res is None:
        res = ref.cw_select(rows, mode, f, scales=scales, virtual=virtual)
        if out is not None:
            out.copy_(res)
            res = out
    if update is not None:
        ref.sgd_step(res, **update)
    return res


_CW_IMPL = {"auto": 0, "direct": 1, "staged": 2, "tiled": 3}


def _cw_impl(impl: str) -> str:
    if impl not in _CW_IMPL:
        raise ValueError(f"impl must be one of {sorted(_CW_IMPL)} (got {impl!r})")
    if impl == "auto":
        env = os.environ.get("BYZPY_CW_IMPL", "auto").lower()
        return env if env in _CW_IMPL else "auto"
    return impl


HOST_MAX_ROWS = 1024


def _host_cw_select(rows: List[torch.Tensor], mode: int, f: int, scales, virtual, out) -> Optional[torch.Tensor]:
    """CPU fp32 rows: the native selection-network kernel (csrc/host_select.cpp) -- tiles of the rows are
    sorted in cache by packed min/max, nothing is stacked.  Returns None when it does not apply (other
    dtypes / devices, synthesised rows, extension not built), leaving the PyTorch implementation."""
    if virtual is not None and virtual[0] > 0:
        return None
    n = len(rows)
    if n > HOST_MAX_ROWS or any(r.device.type != "cpu" or r.dtype != torch.float32 for r in rows):
        return None
    if out is not None and (out.device.type != "cpu" or out.dtype != torch.float32 or not out.is_contiguous()
                            or out.numel() != rows[0].numel()):
        return None
    ext = _load_ext()
    if ext is None or not hasattr(ext, "host_cw_select"):
        return None
    rows = _prep(rows)
    d = rows[0].numel()
    res = out if out is not None else torch.empty(d, dtype=torch.float32)
    ext.host_cw_select([r.data_ptr() for r in rows], _scales(scales, n), int(mode), int(f), d,
                       res.data_ptr(), torch.get_num_threads())
    return res


def _unpack_update(update: Optional[dict]):
    if not update:
        return [], [], 0.0, 0.0, 0.0
    params = list(update["params"])
    moms = list(update.get("moms") or [])
    return (params, moms, float(update.get("lr", 0.0)), float(update.get("momentum", 0.0)),
            float(update.get("weight_decay", 0.0)))


def cw_median(rows: Rows, **kw) -> torch.Tensor:
    """Coordinate-wise (lower) median of the rows; keyword arguments as for :func:`cw_select`."""
    return cw_select(rows, MODE_MEDIAN, 0, **kw)


def cw_trimmed_mean(rows: Rows, f: int, **kw) -> torch.Tensor:
    """Coordinate-wise mean after trimming the ``f`` smallest and ``f`` largest values; see :func:`cw_select`."""
    return cw_select(rows, MODE_TRMEAN, f, **kw)


def cw_meamed(rows: Rows, f: int, **kw) -> torch.Tensor:
    """Coordinate-wise mean of the ``n - f`` values closest to the median; see :func:`cw_select`."""
    return cw_select(rows, MODE_MEAMED, f, **kw)


def cw_mean(rows: Rows, **kw) -> torch.Tensor:
    """Coordinate-wise mean through the selection kernel's load path (same row handling: scales, synthesised rows); see
    :func:`cw_select`."""
    return cw_select(rows, MODE_MEAN, 0, **kw)


# ----------------------------------------------------------------------------
# Gram family
# ----------------------------------------------------------------------------
_GRAM_SCRATCH: dict = {}


def _gram_scratch(dev: torch.device, n: int) -> torch.Tensor:
    ext = require_ext()
    need = ext.gram_partials_needed(n, sm_count(dev))
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _GRAM_SCRATCH.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.float32, device=dev)
        _GRAM_SCRATCH[key] = buf
    return buf


def gram_with_median(rows: Rows, *, scales: Optional[Sequence[float]] = None,
                     want64: bool = False) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """One pass over the rows that produces BOTH the coordinate-wise lower median ``m`` of the (scaled)
    rows and the Gram matrix of ``[rows..., m]`` (``(n+1, n+1)``): the start point of Weiszfeld /
    centred clipping costs no extra pass over the n x d matrix.  Returns ``None`` when the fused
    kernel does not apply (CPU tensors, n > 16) -- callers then run the two passes."""
    rows = as_rows(rows)
    n = len(rows)
    if not _kernel_ok(rows) or n > 16:
        return None
    ext = require_ext()
    rows = _prep(rows)
    d = rows[0].numel()
    dev = rows[0].device
    med = torch.empty(d, dtype=torch.float32, device=dev)
    G = torch.empty((n + 1, n + 1), dtype=torch.float32, device=dev)
    G64 = torch.empty((n + 1, n + 1), dtype=torch.float64, device=dev) if want64 else None
    scratch = _gram_scratch(dev, n + 1)
    ext.gram([r.data_ptr() for r in rows], _scales(scales, n), 0, d, scratch.data_ptr(),
             scratch.numel() // ((n + 1) * (n + 1)), G.data_ptr(), G64.data_ptr() if want64 else 0,
             sm_count(dev), _stream(dev), med.data_ptr())
    return (G64 if want64 else G), med


def gram(rows: Rows, *, scales: Optional[Sequence[float]] = None, want64: bool = False,
         impl: str = "auto", diag_only: bool = False) -> torch.Tensor:
    """``G = (S X)(S X)^T`` as an ``(n, n)`` tensor on the rows' device.

    ``impl``: ``"auto"`` | ``"fp32"`` (CUDA-core exact) | ``"umma"`` (tcgen05 3xTF32).
    ``diag_only``: the caller only uses the squared row norms (the torch fallback then skips the
    off-diagonal work; the CUDA kernels are bandwidth bound and compute the full matrix anyway).
    """
    rows = as_rows(rows)
    n = len(rows)
    d = rows[0].numel()
    if _kernel_ok(rows):
        ext = require_ext()
        rows = _prep(rows)
        dev = rows[0].device
        use_umma = impl == "umma" or (impl == "auto" and hasattr(ext, "gram_umma") and n > 16
                                      and d >= 4096)
        G = torch.empty((n, n), dtype=torch.float32, device=dev)
        G64 = torch.empty((n, n), dtype=torch.float64, device=dev) if want64 else None
        if use_umma:
            from . import umma

            umma.gram_umma(rows, _scales(scales, n), G, G64)
        else:
            scratch = _gram_scratch(dev, n)
            ext.gram(
                [r.data_ptr() for r in rows], _scales(scales, n), 0, d, scratch.data_ptr(),
                scratch.numel() // (n * n), G.data_ptr(), G64.data_ptr() if want64 else 0,
                sm_count(dev), _stream(dev),
            )
        return G64 if want64 else G
    return ref.gram(rows, scales=scales, want64=want64, diag_only=diag_only)


def sqdist_from_gram(G: torch.Tensor) -> torch.Tensor:
    """Pairwise squared distances ``D_ij = G_ii + G_jj - 2 G_ij`` clamped at 0."""
    diag = torch.diagonal(G)
    D = diag[:, None] + diag[None, :] - 2.0 * G
    D = torch.clamp(D, min=0.0)
    D = torch.nan_to_num(D, nan=float("inf"))
    D.fill_diagonal_(0.0)
    return D


def weighted_sum(
    rows: Rows,
    W: torch.Tensor,
    *,
    scales: Optional[Sequence[float]] = None,
    out: Optional[torch.Tensor] = None,
    update: Optional[dict] = None,
    multi_impl: str = "auto",
) -> torch.Tensor:
    """``Y = W (S X)`` with ``W`` an ``(m, n)`` (or ``(n,)``) weight tensor on the device.

    Returns ``(m, d)`` (or ``(d,)`` for 1-D ``W``).  ``m <= 8``: the streaming pointer-table kernel
    (one pass, zero-weight rows skipped).  ``m > 8``: one pass too -- a register-tiled fp32 GEMM over
    shared-memory tiles of the inputs (``csrc/wsum.cu``, ``wsum_multi``) -- unless ``W`` is sparse enough
    that the 8-rows-per-pass kernel reads less (a diagonal map reads every input once either way).
    """
    rows = as_rows(rows)
    n = len(rows)
    d = rows[0].numel()
    squeeze = W.dim() == 1
    W2 = W.reshape(1, -1) if squeeze else W
    m = W2.shape[0]
    if W2.shape[1] != n:
        raise ValueError("W must have one column per row")
    if _kernel_ok(rows):
        ext = require_ext()
        rows = _prep(rows)
        dev = rows[0].device
        Wd = W2.to(device=dev, dtype=torch.float32).contiguous()
        if out is None:
            out2 = torch.empty((m, d), dtype=torch.float32, device=dev)
        else:
            out2 = out.reshape(m, d)
        params, moms, lr, mu, wd = _unpack_update(update)
        ptrs = [r.data_ptr() for r in rows]
        sc = _scales(scales, n)
        d_main = 0
        if m > 8 and not params and hasattr(ext, "wsum_multi") and _wsum_dense(W2, multi_impl):
            tile = ext.WSUM_MULTI_TILE
            d_main = (d // tile) * tile
            if d_main and all(p % 16 == 0 for p in ptrs) and out2.data_ptr() % 16 == 0 and (d * 4) % 16 == 0:
                ext.wsum_multi(ptrs, sc, Wd.data_ptr(), m, 0, d_main, [out2[r].data_ptr() for r in range(m)],
                               sm_count(dev), _stream(dev))
            else:
                d_main = 0
        if d_main == d:
            return out2[0] if squeeze else out2
        # the streaming kernel emits up to 8 output rows per pass over the inputs (here: the remaining columns)
        if d_main:
            ptrs = [p + 4 * d_main for p in ptrs]
        for r0 in range(0, m, 8):
            mb = min(8, m - r0)
            first = r0 == 0
            ext.wsum(
                ptrs, sc, Wd[r0:r0 + mb].data_ptr(), mb, 0, d - d_main,
                [out2[r0 + r].data_ptr() + 4 * d_main for r in range(mb)],
                [p.data_ptr() for p in params] if first else [],
                [mm.data_ptr() for mm in moms] if first else [],
                lr, mu, wd, sm_count(dev), _stream(dev),
            )
        return out2[0] if squeeze else out2
    res = ref.weighted_sum(rows, W2, scales=scales)
    if out is not None:
        out.reshape(m, d).copy_(res)
        res = out.reshape(m, d)
    if update is not None:
        ref.sgd_step(res[0], **update)
    return res[0] if squeeze else res


def _wsum_dense(W: torch.Tensor, impl: str) -> bool:
    """Use the one-pass multi-row kernel?  ``"multi"`` / ``"passes"`` force one form (benchmarks, tests);
    ``"auto"`` compares the two cost models without looking at the (device-resident) weights: the
    8-rows-per-pass kernel moves ``ceil(m / 8) n + m`` floats per coordinate at the HBM rate, the
    one-pass kernel issues ``m n`` FMAs per coordinate at about half the fp32 FMA peak (measured)."""
    if impl == "passes":
        return False
    if impl == "multi":
        return True
    m, n = W.shape
    passes_s = ((m + 7) // 8 * n + m) * 4 / 6.4e12
    # measured fp32-FMA efficiency by register-tile height (profiles/wsum_multi.md): 13 TFMA/s at m = 32,
    # 21 TFMA/s at m = 64 / 128 of the 37 TFMA/s peak
    eff = 0.2 if m <= 16 else (0.35 if m <= 32 else (0.42 if m <= 64 else 0.55))
    multi_s = m * n / (eff * 37.2e12)
    return multi_s < passes_s


def colstat(rows: Rows, a: float, b: float, *, scales=None, out=None) -> torch.Tensor:
    """``a*mean + b*std`` (population std) per coordinate."""
    rows = as_rows(rows)
    n = len(rows)
    d = rows[0].numel()
    if _kernel_ok(rows):
        ext = require_ext()
        rows = _prep(rows)
        dev = rows[0].device
        if out is None:
            out = torch.empty(d, dtype=torch.float32, device=dev)
        ext.colstat([r.data_ptr() for r in rows], _scales(scales, n), float(a), float(b), 0, d,
                    out.data_ptr(), sm_count(dev), _stream(dev))
        return out
    ext = _load_ext()
    if (ext is not None and hasattr(ext, "host_colstat")
            and all(r.device.type == "cpu" and r.dtype == torch.float32 for r in rows)
            and (out is None or (out.device.type == "cpu" and out.dtype == torch.float32
                                 and out.is_contiguous() and out.numel() == d))):
        # native two-sweep column statistic over the row pointers (csrc/host_select.cpp): no (n, d) stack
        rows = _prep(rows)
        res = out if out is not None else torch.empty(d, dtype=torch.float32)
        ext.host_colstat([r.data_ptr() for r in rows], _scales(scales, n), float(a), float(b), d,
                         res.data_ptr(), torch.get_num_threads())
        return res
    res = ref.colstat(rows, a, b, scales=scales)
    if out is not None:
        out.copy_(res)
        return out
    return res


# ----------------------------------------------------------------------------
# element-wise helpers
# ----------------------------------------------------------------------------
def scale_copy(src: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out = scale * src`` in one pass (a grid-stride kernel for fp32 CUDA tensors, PyTorch otherwise); ``out`` is
    allocated when not given.
    """
    flat = src.reshape(-1)
    if flat.is_cuda and flat.dtype == torch.float32:
        ext = require_ext()
        flat = flat.contiguous()
        if out is None:
            out = torch.empty_like(flat)
        ext.scale_copy(flat.data_ptr(), out.data_ptr(), float(scale), flat.numel(),
                       sm_count(flat.device), _stream(flat.device))
        return out.reshape(src.shape)
    res = src * scale
    if out is not None:
        out.copy_(res.reshape(out.shape))
        return out
    return res


def fill_(dst: torch.Tensor, value: float) -> torch.Tensor:
    """Fill ``dst`` with ``value`` in place (kernel for contiguous fp32 CUDA tensors) and return it."""
    if dst.is_cuda and dst.dtype == torch.float32 and dst.is_contiguous():
        ext = require_ext()
        ext.fill(dst.data_ptr(), float(value), dst.numel(), sm_count(dst.device),
                 _stream(dst.device))
        return dst
    return dst.fill_(value)


def gaussian_(dst: torch.Tensor, mu: float, sigma: float, seed: int, offset: int = 0) -> torch.Tensor:
    """Fill ``dst`` with N(mu, sigma^2) samples (Philox4x32-10 counter RNG on CUDA)."""
    if dst.is_cuda and dst.dtype == torch.float32 and dst.is_contiguous():
        ext = require_ext()
        ext.gaussian(dst.data_ptr(), float(mu), float(sigma), int(seed) & (2**64 - 1),
                     int(offset), dst.numel(), sm_count(dst.device), _stream(dst.device))
        return dst
    gen = torch.Generator(device="cpu")
    gen.manual_seed(int(seed) & (2**63 - 1))
    vals = torch.randn(dst.numel() + offset, generator=gen, dtype=torch.float32)[offset:]
    dst.copy_((vals * sigma + mu).reshape(dst.shape).to(dst.dtype))
    return dst


def normalize_uint8_nhwc(x: torch.Tensor, mean: Union[float, Sequence[float]] = 127.5,
                         std: Union[float, Sequence[float]] = 127.5, *, s2d: bool = False):
    """uint8 image batch ``[N, H, W, C]`` -> ``(x - mean[c]) / std[c]`` as a bf16 ``[N, C, H, W]`` tensor
    with channels-last strides (the same memory order, no transpose): the whole input pipeline of a
    replica step in one streaming kernel instead of float() / sub / mul / cast passes."""
    if x.dtype != torch.uint8 or x.dim() != 4:
        raise TypeError("expected a uint8 [N, H, W, C] tensor")
    C = x.shape[3]
    mean = [float(mean)] * C if not isinstance(mean, (list, tuple)) else [float(v) for v in mean]
    std = [float(std)] * C if not isinstance(std, (list, tuple)) else [float(v) for v in std]
    if len(mean) != C or len(std) != C:
        raise ValueError("mean / std must have one entry per channel")
    if s2d and x.is_cuda and x.is_contiguous() and C == 3:
        # ``s2d=True``: emit the 2x2 space-to-depth packed, zero-padded form consumed by
        # ``ops.fused_layers.S2DStemConv2d`` (the ResNet stem) -- same single pass over the bytes
        from .fused_layers import pack_stem_input

        return pack_stem_input(x, mean, std)
    scale = [1.0 / v for v in std]
    if x.is_cuda and x.is_contiguous() and C <= 8 and x.data_ptr() % 16 == 0:
        ext = require_ext()
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        ext.u8_affine(x.data_ptr(), out.data_ptr(), x.numel(), C, mean, scale, sm_count(x.device),
                      _stream(x.device))
        count_launch()
        return out.permute(0, 3, 1, 2)
    m = torch.tensor(mean, dtype=torch.float32, device=x.device)
    s = torch.tensor(scale, dtype=torch.float32, device=x.device)
    return ((x.float() - m) * s).to(torch.bfloat16).permute(0, 3, 1, 2)


def sgd_step(grad: torch.Tensor, params: Sequence[torch.Tensor],
             moms: Optional[Sequence[torch.Tensor]] = None, *, lr: float, momentum: float = 0.0,
             weight_decay: float = 0.0) -> None:
    """Fused SGD(+momentum) over flat replicas: one kernel for all of them."""
    params = list(params)
    moms = list(moms) if moms else []
    if grad.is_cuda and grad.dtype == torch.float32 and len(params) <= MAXR:
        ext = require_ext()
        ext.sgd(grad.data_ptr(), [p.data_ptr() for p in params], [m.data_ptr() for m in moms],
                float(lr), float(momentum), float(weight_decay), grad.numel(),
                sm_count(grad.device), _stream(grad.device))
        return
    ref.sgd_step(grad, params=params, moms=moms or None, lr=lr, momentum=momentum,
                 weight_decay=weight_decay)


__all__ = [
    "MODE_MEDIAN", "MODE_TRMEAN", "MODE_MEAMED", "MODE_MEAN", "as_rows", "cw_select", "cw_median",
    "cw_trimmed_mean", "cw_meamed", "cw_mean", "gram", "gram_with_median", "sqdist_from_gram", "weighted_sum",
    "colstat", "scale_copy", "fill_", "gaussian_", "sgd_step", "extension_available",
    "require_ext", "sm_count", "normalize_uint8_nhwc", "count_launch", "launches",
]
