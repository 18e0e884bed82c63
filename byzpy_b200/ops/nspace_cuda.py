"""Device-side n-space solvers (single-CTA CUDA kernels in ``csrc/nspace.cu``): the selection /
iteration logic of reference krum.py:177-194, geometric_median.py:87-102, center_clipping.py:146-154,
minimum_diameter_average.py:358-386 and smea.py:63-88, evaluated on the (n, n) Gram matrix.

Each function takes the fp64 Gram matrix ON THE DEVICE and returns the weight /
coefficient vector ON THE DEVICE without any host synchronisation, so a
Gram-family aggregation is three back-to-back launches (gram -> solve ->
weighted sum).  Returns ``None`` when the extension lacks the kernel, in which
case the caller falls back to the host solver in :mod:`byzpy_b200.ops.nspace`.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _load_ext, _stream


_CONST_CACHE: dict = {}


def _device_const(a0: np.ndarray, device: torch.device) -> torch.Tensor:
    """Start-coefficient vectors are tiny constants: upload once per (device, value) so that no
    host->device copy happens inside a CUDA-graph capture."""
    arr = np.ascontiguousarray(np.asarray(a0, dtype=np.float64))
    key = (str(device), arr.tobytes())
    t = _CONST_CACHE.get(key)
    if t is None:
        if len(_CONST_CACHE) > 256:
            _CONST_CACHE.clear()
        t = torch.from_numpy(arr.copy()).to(device)
        _CONST_CACHE[key] = t
    return t


def _ready(G: torch.Tensor, name: str):
    ext = _load_ext()
    if ext is None or not hasattr(ext, name) or not G.is_cuda or G.shape[0] > 128:
        return None
    return ext


def krum_weights(G: torch.Tensor, f: int, q: int) -> Optional[torch.Tensor]:
    """Multi-Krum weights from the device Gram matrix, by one single-CTA kernel (no host round trip); ``None`` when the
    kernel library lacks it or ``n > 128``.
    """
    ext = _ready(G, "nspace_krum")
    if ext is None:
        return None
    n = G.shape[0]
    G = G.contiguous().double()
    w = torch.empty(n, dtype=torch.float32, device=G.device)
    ext.nspace_krum(G.data_ptr(), n, int(f), int(q), w.data_ptr(), _stream(G.device))
    return w


def weiszfeld_coeffs(G: torch.Tensor, n_real: int, a0: np.ndarray, *, tol: float, max_iter: int,
                     eps: float) -> Optional[torch.Tensor]:
    """Coefficients of the geometric median over ``[rows; start point]`` from the device Gram matrix: the whole Weiszfeld
    iteration runs inside one single-CTA kernel (fp64); ``None`` when unavailable.
    """
    ext = _ready(G, "nspace_weiszfeld")
    if ext is None:
        return None
    nt = G.shape[0]
    G = G.contiguous().double()
    a = _device_const(a0, G.device)
    w = torch.empty(nt, dtype=torch.float32, device=G.device)
    iters = torch.zeros(1, dtype=torch.int32, device=G.device)
    ext.nspace_weiszfeld(G.data_ptr(), nt, int(n_real), a.data_ptr(), float(tol), int(max_iter),
                         float(eps), w.data_ptr(), iters.data_ptr(), _stream(G.device))
    return w


def centered_clip_coeffs(G: torch.Tensor, n_real: int, a0: np.ndarray, *, c_tau: float, M: int,
                         eps: float) -> Optional[torch.Tensor]:
    """Coefficients of ``M`` centered-clipping rounds from the device Gram matrix (one single-CTA kernel, fp64);
    ``None`` when unavailable.
    """
    ext = _ready(G, "nspace_cclip")
    if ext is None:
        return None
    nt = G.shape[0]
    G = G.contiguous().double()
    a = _device_const(a0, G.device)
    w = torch.empty(nt, dtype=torch.float32, device=G.device)
    ext.nspace_cclip(G.data_ptr(), nt, int(n_real), a.data_ptr(), float(c_tau), int(M), float(eps),
                     w.data_ptr(), _stream(G.device))
    return w


def _uniform_on_smallest(scores: torch.Tensor, k: int) -> torch.Tensor:
    """fp32 weights 1/k on the k smallest entries of ``scores`` (ties to the lower index), built from
    sort + index_fill only: no host synchronisation, so the round stays CUDA-graph capturable."""
    order = torch.sort(scores, stable=True).indices[:k]
    w = torch.zeros(scores.shape[0], dtype=torch.float32, device=scores.device)
    return w.index_fill_(0, order, 1.0 / k)


def cge_weights(G: torch.Tensor, n: int, f: int) -> torch.Tensor:
    """Drop the f largest-norm rows among the first ``n`` (``ops.nspace.cge_weights`` on the device)."""
    g = torch.diagonal(G)[:n].double().clamp_min(0.0)
    g = torch.where(torch.isnan(g), torch.full_like(g, float("inf")), g)
    w = torch.zeros(G.shape[0], dtype=torch.float32, device=G.device)
    w[:n] = _uniform_on_smallest(g, n - f)            # norms and squared norms order identically
    return w


def monna_weights(G: torch.Tensor, n: int, f: int, reference_index: int) -> torch.Tensor:
    """n-f rows nearest to the trusted row (``ops.nspace.monna_weights`` on the device)."""
    G = G.double()
    g = torch.diagonal(G)[:n]
    d = g + g[reference_index] - 2.0 * G[reference_index, :n]
    d = torch.where(torch.isnan(d), torch.full_like(d, float("inf")), d).clamp_min(0.0)
    # the trusted row always comes first; written as a select (no scalar-into-slice assignment: that path
    # materialises the Python scalar through a host tensor, which a capturing stream rejects)
    trusted = torch.arange(n, device=d.device) == int(reference_index)
    d = torch.where(trusted, torch.full_like(d, -1.0), d)
    w = torch.zeros(G.shape[0], dtype=torch.float32, device=G.device)
    w[:n] = _uniform_on_smallest(d, n - f)
    return w


def _preagg(G: torch.Tensor, n: int, mode: int, param: float, iparam: int, want32: bool):
    ext = _ready(G, "nspace_preagg")
    if ext is None or G.shape[0] != n or G.shape[1] != n:
        return None
    G = G.contiguous().double()
    W = torch.empty((n, n), dtype=torch.float64, device=G.device)
    W32 = torch.empty((n, n), dtype=torch.float32, device=G.device) if want32 else None
    ext.nspace_preagg(G.data_ptr(), n, mode, float(param), int(iparam), W.data_ptr(),
                      W32.data_ptr() if W32 is not None else 0, _stream(G.device))
    return (W, W32) if want32 else W


def clip_matrix(G: torch.Tensor, threshold: float, *, want32: bool = False):
    """``diag(min(1, threshold / ||x_i||))`` as an (n, n) device matrix (``ops.nspace.clip_scales``)."""
    return _preagg(G, G.shape[0], 0, threshold, 0, want32)


def arc_matrix(G: torch.Tensor, f: int, *, want32: bool = False):
    """Adaptive robust clipping map (``ops.nspace.arc_scales``) on the device."""
    return _preagg(G, G.shape[0], 1, 0.0, f, want32)


def nnm_matrix(G: torch.Tensor, f: int, *, want32: bool = False):
    """Nearest-neighbour mixing matrix (``ops.nspace.nnm_matrix``) on the device."""
    return _preagg(G, G.shape[0], 2, 0.0, f, want32)


def caf_coeffs(G: torch.Tensor, n: int, f: int, *, power_iters: int) -> Optional[torch.Tensor]:
    """CAF weights over the n rows (+ a zero for the start-direction row) from the (n+1, n+1) Gram
    of ``[rows; r]`` (``ops.nspace.caf_coeffs``) -- the whole filter loop in one single-CTA kernel."""
    ext = _load_ext()
    if ext is None or not hasattr(ext, "nspace_caf") or not G.is_cuda or n > 127 or G.shape[0] != n + 1:
        return None
    G = G.contiguous().double()
    w = torch.empty(n + 1, dtype=torch.float32, device=G.device)
    ext.nspace_caf(G.data_ptr(), int(n), int(f), int(power_iters), w.data_ptr(), _stream(G.device))
    return w


__all__ = ["krum_weights", "weiszfeld_coeffs", "centered_clip_coeffs", "cge_weights", "monna_weights",
           "clip_matrix", "arc_matrix", "nnm_matrix", "caf_coeffs"]


SUBSET_MAX_N = 24
SUBSET_MAX_COMBOS = 4_000_000


def subset_search_feasible(n: int, m: int) -> bool:
    """True when the exhaustive device search handles (n, m): n <= 24 and C(n, m) <= 4 M subsets."""
    ext = _load_ext()
    if ext is None or not hasattr(ext, "nspace_subset") or not (1 <= m <= n <= SUBSET_MAX_N):
        return False
    return ext.binomial(n, m) <= SUBSET_MAX_COMBOS


def subset_weights(G: torch.Tensor, n: int, m: int, mode: str) -> Optional[torch.Tensor]:
    """1/m on the rows of the optimal m-subset of the first ``n`` rows of the fp64 Gram ``G``.
    ``mode``: ``"mda"`` (minimum diameter) or ``"smea"`` (minimum top covariance eigenvalue); ties go
    to the lexicographically first subset, like ``ops.nspace.mda_subset`` / ``smea_subset``."""
    if not G.is_cuda or not subset_search_feasible(n, m):
        return None
    from . import sm_count

    ext = _load_ext()
    G = G.contiguous().double()
    nt = G.shape[0]
    dev = G.device
    sms = sm_count(dev)
    nb = ext.nspace_subset_blocks(n, m, sms)
    score = torch.empty(nb, dtype=torch.float64, device=dev)
    rank = torch.empty(nb, dtype=torch.int64, device=dev)
    w = torch.empty(nt, dtype=torch.float32, device=dev)
    ext.nspace_subset(G.data_ptr(), nt, n, m, nt, 0 if mode == "mda" else 1, score.data_ptr(), rank.data_ptr(),
                      w.data_ptr(), sms, _stream(dev))
    return w
