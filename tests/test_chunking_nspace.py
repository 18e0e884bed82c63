import itertools

import numpy as np
import pytest

from byzpy_b200.aggregators._chunking import select_adaptive_chunk_size
from byzpy_b200.ops import nspace


def test_chunking_basic_and_degenerate():
    assert select_adaptive_chunk_size(0, 10) == 0
    assert select_adaptive_chunk_size(5, 10) == 5
    assert select_adaptive_chunk_size(100, 32, pool_size=1) == 32
    assert select_adaptive_chunk_size(65536, 8192, pool_size=6) == 2731
    assert select_adaptive_chunk_size(65536, 8192, pool_size=2) == 8192
    assert select_adaptive_chunk_size(1000, 100, pool_size=4) == 63


def test_chunking_env_overrides(monkeypatch):
    base = select_adaptive_chunk_size(10**6, 4096, pool_size=6)
    monkeypatch.setenv("BYZPY_CHUNK_MIN_PER_WORKER", "64")
    more = select_adaptive_chunk_size(10**6, 4096, pool_size=6)
    assert more < base
    monkeypatch.setenv("BYZPY_CHUNK_MAX_SHRINK", "2")
    assert select_adaptive_chunk_size(10**6, 4096, pool_size=6) >= 2048
    monkeypatch.setenv("BYZPY_CHUNK_TARGET_FACTOR", "not-a-number")
    select_adaptive_chunk_size(10**6, 4096, pool_size=6)  # bad values are ignored


def _X(n, d, seed):
    return np.random.default_rng(seed).normal(size=(n, d))


def test_sqdist_and_krum_bruteforce():
    X = _X(9, 20, 0)
    G = X @ X.T
    D = nspace.sqdist(G)
    ref = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    assert np.allclose(D, ref, atol=1e-9)
    f, q = 2, 3
    scores = np.array([np.sort(np.delete(ref[i], i))[: 9 - f - 1].sum() for i in range(9)])
    assert np.allclose(nspace.krum_scores(G, f), scores)
    w = nspace.krum_weights(G, f, q)
    assert set(np.nonzero(w)[0]) == set(np.argsort(scores)[:q]) and np.isclose(w.sum(), 1.0)


def test_mda_subset_is_lexicographically_first_minimiser():
    for seed in range(5):
        X = _X(8, 3, seed)
        D = nspace.sqdist(X @ X.T)
        m = 5
        best, best_combo = None, None
        for combo in itertools.combinations(range(8), m):
            diam = max(D[i, j] for i in combo for j in combo)
            if best is None or diam < best:
                best, best_combo = diam, combo
        assert nspace.mda_subset(D, m) == best_combo


def test_mda_tie_breaks_to_first_subset():
    D = np.ones((5, 5)) - np.eye(5)
    assert nspace.mda_subset(D, 3) == (0, 1, 2)


def test_smea_subset_bruteforce():
    X = _X(7, 10, 4)
    G = X @ X.T
    m = 5
    best, best_combo = None, None
    for combo in itertools.combinations(range(7), m):
        sub = X[list(combo)]
        cov = np.cov(sub.T, bias=True)
        lam = np.linalg.eigvalsh(cov)[-1]
        if best is None or lam < best - 1e-12:
            best, best_combo = lam, combo
    assert nspace.smea_subset(G, m) == best_combo


def test_weiszfeld_coeffs_match_data_space_iteration():
    X = _X(8, 30, 7)
    G = X @ X.T
    a0 = np.full(8, 1 / 8)
    a, iters = nspace.weiszfeld_coeffs(G, 8, a0, tol=1e-10, max_iter=300, eps=1e-12)
    z = X.mean(0)
    for _ in range(300):
        dist = np.maximum(np.linalg.norm(X - z, axis=1), 1e-12)
        w = 1 / dist
        z_new = (w[:, None] * X).sum(0) / w.sum()
        if np.linalg.norm(z_new - z) <= 1e-10:
            z = z_new
            break
        z = z_new
    assert np.allclose(a @ X, z, atol=1e-7) and 1 <= iters <= 300


def test_centered_clip_coeffs_match_data_space():
    X = _X(6, 12, 9)
    G = X @ X.T
    a = nspace.centered_clip_coeffs(G, 6, np.zeros(6), c_tau=0.8, M=5, eps=1e-12)
    v = np.zeros(12)
    for _ in range(5):
        diff = X - v
        dist = np.maximum(np.linalg.norm(diff, axis=1), 1e-12)
        v = v + (np.minimum(1.0, 0.8 / dist)[:, None] * diff).sum(0) / 6
    assert np.allclose(a @ X, v, atol=1e-9)


def test_selection_validation():
    G = np.eye(4)
    with pytest.raises(ValueError):
        nspace.krum_weights(G, 3, 1)
    with pytest.raises(ValueError):
        nspace.monna_weights(G, 2)
    with pytest.raises(ValueError):
        nspace.cge_weights(G, 4)
    with pytest.raises(ValueError):
        nspace.bucket_matrix(4, 2, [0, 1, 2, 2])


# ---------------------------------------------------------------- torch fallback fast paths
def test_reference_gram_variants_agree():
    import torch

    from byzpy_b200.ops import reference as ref

    torch.manual_seed(0)
    rows = [torch.randn(20_000) for _ in range(7)]            # > one 8192-column chunk
    X = torch.stack(rows).double()
    exact = X @ X.T
    G = ref.gram(rows, want64=True)
    assert G.dtype == torch.float64
    torch.testing.assert_close(G, exact, rtol=1e-6, atol=1e-3)
    D = ref.gram(rows, want64=True, diag_only=True)
    # squared norms come from BLAS fp32 dots (1e-7 relative), the same band as the full Gram above
    torch.testing.assert_close(torch.diagonal(D), torch.diagonal(exact), rtol=1e-6, atol=1e-3)
    assert float((D - torch.diag(torch.diagonal(D))).abs().max()) == 0.0
    Ds = ref.gram(rows, want64=True, diag_only=True, scales=[2.0] * 7)
    torch.testing.assert_close(torch.diagonal(Ds), 4.0 * torch.diagonal(exact), rtol=1e-6, atol=1e-3)
    big = [torch.full((1000,), 1e20), torch.tensor([float("inf"), 1.0]).repeat(500), torch.randn(1000).double()]
    Db = torch.diagonal(ref.gram(big, want64=True, diag_only=True))
    assert float(Db[0]) == pytest.approx(1e43, rel=1e-6)      # fp32 squares overflow: fp64 norm takes over
    assert float(Db[1]) == float("inf")
    assert float(Db[2]) == pytest.approx(float(big[2] @ big[2]), rel=1e-12)


def test_reference_weighted_sum_sparse_dense_and_single_row_paths():
    import torch

    from byzpy_b200.ops import reference as ref

    torch.manual_seed(1)
    n, d = 32, 500
    rows = [torch.randn(d) for _ in range(n)]
    X = torch.stack(rows)
    W_sparse = torch.zeros(4, n)                               # bucket means: 4 non-zeros per row
    for r in range(4):
        W_sparse[r, 8 * r: 8 * r + 4] = 0.25
    torch.testing.assert_close(ref.weighted_sum(rows, W_sparse), W_sparse @ X)
    W_dense = torch.rand(5, n)
    torch.testing.assert_close(ref.weighted_sum(rows, W_dense), W_dense @ X, rtol=1e-5, atol=1e-5)
    W_diag = torch.diag(torch.rand(n))
    torch.testing.assert_close(ref.weighted_sum(rows, W_diag), W_diag @ X, rtol=1e-6, atol=1e-6)
    w1 = torch.rand(n)
    w1[3] = 0.0
    rows_inf = list(rows)
    rows_inf[3] = torch.full((d,), float("inf"))               # zero weight: never touched
    out = ref.weighted_sum(rows_inf, w1)
    assert out.shape == (1, d) and torch.isfinite(out).all()
    torch.testing.assert_close(out[0], (w1[:, None] * torch.stack(rows)).sum(0) - w1[3] * rows[3], rtol=1e-5,
                               atol=1e-5)


def test_stack_cache_is_invalidated_by_in_place_updates():
    import torch

    from byzpy_b200.ops import reference as ref

    rows = [torch.ones(9000) * (i + 1) for i in range(3)]
    ref.gram(rows)                                             # remembers the (3, d) stack
    rows[0].mul_(10.0)                                         # in-place change bumps the version counter
    W = torch.eye(3)[:2] + 0.5                                 # dense 2 x 3 -> the stacked path
    out = ref.weighted_sum(rows, W)
    torch.testing.assert_close(out, W @ torch.stack(rows))


def test_stack_cache_rejects_rows_whose_ids_were_recycled(monkeypatch):
    import gc

    import torch

    from byzpy_b200.ops import reference as ref

    old = [torch.ones(9000) * (i + 1) for i in range(3)]
    ref.gram(old)                                              # remembers the stack of `old`
    del old
    gc.collect()
    # worst case of id reuse: the key of the new rows collides with the remembered one
    monkeypatch.setattr(ref, "_stack_key", lambda rows, scales: "same-key")
    ref._STACK_CACHE["entry"] = ("same-key",) + ref._STACK_CACHE["entry"][1:]
    new = [torch.full((9000,), -1.0) * (i + 1) for i in range(3)]
    W = torch.eye(3)[:2] + 0.5
    torch.testing.assert_close(ref.weighted_sum(new, W), W @ torch.stack(new))
    # ... while the very same live tensors do hit
    ref.gram(new)
    cached = ref._STACK_CACHE["entry"][1]
    assert ref._stack(new, None, reuse=True) is cached and "entry" not in ref._STACK_CACHE
