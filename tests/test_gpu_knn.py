"""GPU: the cosine kNN builder (next row f1) against the oracle's restatement of graphs/build_ii_knn.py."""
import numpy as np
import pytest
import torch

from oracle import gat_oracle as O

pytestmark = pytest.mark.gpu


def _data(n, seed, clustered, d=128):
    rng = np.random.default_rng(seed)
    if clustered:
        centers = rng.standard_normal((max(n // 40, 2), d))
        x = centers[rng.integers(0, len(centers), n)] + 0.6 * rng.standard_normal((n, d))
    else:
        x = rng.standard_normal((n, d))
    return (x * rng.uniform(0.5, 3.0, size=(n, 1))).astype(np.float32)     # un-normalised rows, like raw embeddings


@pytest.mark.parametrize("n,clustered,k,min_sim,d", [(50, False, 20, -1.0, 128), (129, True, 20, 0.3, 128),
                                                     (3000, True, 20, 0.3, 128), (2049, False, 5, -1.0, 128),
                                                     (5000, True, 20, 0.5, 128),
                                                     # >= 64 column blocks: the append pipeline (group maxima -> sampled appends -> full sweep)
                                                     (9000, True, 20, 0.3, 128), (12345, False, 20, -1.0, 128),
                                                     (20000, True, 32, 0.3, 128),
                                                     # the 384-d text embeddings (embeddings/embed_text.py -> build_ii_knn.py)
                                                     (130, False, 20, -1.0, 384), (3000, True, 20, 0.3, 384),
                                                     (4100, True, 10, 0.5, 384), (8300, True, 20, 0.3, 384)])
def test_knn_matches_oracle(n, clustered, k, min_sim, d, monkeypatch):
    import b200gat
    emb = _data(n, n, clustered, d)
    rows, cols, sims = O.build_ii_knn(emb, k=min(k, n - 1), min_similarity=min_sim, batch_size=1000)
    r, c, s = b200gat.build_ii_knn(torch.from_numpy(emb).cuda(), k=min(k, n - 1), min_similarity=min_sim)
    r, c, s = r.cpu().numpy(), c.cpu().numpy(), s.cpu().numpy()
    assert r.dtype == np.int32 and c.dtype == np.int32 and s.dtype == np.float32
    # exact dense similarities (fp64) to judge near-ties
    en = emb.astype(np.float64)
    en /= np.linalg.norm(en, axis=1, keepdims=True)
    full = en @ en.T
    np.fill_diagonal(full, -np.inf)
    # every emitted pair carries its true similarity (fp32 rounding), rows ascending, similarities descending inside a row
    np.testing.assert_allclose(s, full[r, c], rtol=0, atol=3e-6)
    assert np.all(np.diff(r) >= 0) and np.all((np.diff(s) <= 1e-6) | (np.diff(r) > 0))
    assert np.all(r != c) and np.all(s >= min_sim - 3e-6)
    # same edge count per row as the reference arithmetic, up to pairs that sit within rounding of the k-th value or of
    # the min_similarity cut
    cnt_ref = np.bincount(rows, minlength=n)
    cnt = np.bincount(r, minlength=n)
    kk = min(k, n - 1)
    kth = -np.sort(-full, axis=1)[:, kk - 1]
    near_cut = ((np.abs(full - min_sim) < 5e-6).sum(1) > 0)
    assert np.all((cnt == cnt_ref) | near_cut)
    # the neighbour sets agree except for columns within rounding of the k-th best
    ref_sets = [set() for _ in range(n)]
    for a, b in zip(rows, cols):
        ref_sets[a].add(int(b))
    bad = 0
    for a, b in zip(r, c):
        if int(b) not in ref_sets[a] and abs(full[a, b] - max(kth[a], min_sim)) > 5e-6:
            bad += 1
    assert bad == 0
    assert len(r) >= 0.999 * len(rows)


def test_knn_rejects_unsupported(monkeypatch):
    import b200gat
    with pytest.raises(RuntimeError):
        b200gat.build_ii_knn(torch.randn(10, 128))                 # CPU tensor
    with pytest.raises(RuntimeError, match="128 or 384"):
        b200gat.build_ii_knn(torch.randn(10, 512).cuda())          # widths other than fused (128) / text (384)


@pytest.mark.parametrize("d", [128, 384])
def test_knn_dense_duplicates_take_the_exact_path(d):
    """Rows whose 48 bf16 candidates cannot prove the top-k (clusters of >48 near-identical items) are redone exactly."""
    import b200gat
    rng = np.random.default_rng(5)
    centers = rng.standard_normal((5, d))
    dup = np.repeat(centers, 120, axis=0) + 1e-4 * rng.standard_normal((600, d))          # 5 clumps of 120 near-duplicates
    emb = np.concatenate([dup, rng.standard_normal((424, d))]).astype(np.float32)
    n, k = emb.shape[0], 20
    stats = {}
    idx, sim, counts = b200gat.knn_neighbors(torch.from_numpy(emb).cuda(), k, 0.3, stats=stats)
    assert int(stats["exact_rows"].item()) >= 600
    idx, sim, counts = idx.cpu().numpy(), sim.cpu().numpy(), counts.cpu().numpy()
    en = emb.astype(np.float64)
    en /= np.linalg.norm(en, axis=1, keepdims=True)
    full = en @ en.T
    np.fill_diagonal(full, -np.inf)
    kth = -np.sort(-full, axis=1)[:, k - 1]
    for r in range(n):
        c = idx[r]
        assert len(set(c.tolist())) == k and r not in c
        np.testing.assert_allclose(sim[r], full[r, c], rtol=0, atol=3e-6)
        assert np.all(np.diff(sim[r]) <= 1e-6)
        assert np.all(full[r, c] >= kth[r] - 5e-6)                # every neighbour is a true top-k member up to rounding
    assert np.all(counts[:600] == k)
    rows, cols, sims = O.build_ii_knn(emb, k=k, min_similarity=0.3)
    np.testing.assert_array_equal(np.bincount(rows, minlength=n), counts)


@pytest.mark.parametrize("name", ["knn_128", "knn_384"])
def test_knn_matches_reference_script_output(golden_dir, name):
    """Against the reference itself: fixtures written by graphs/build_ii_knn.py run as a script (oracle/make_golden.py knn)."""
    import os
    import b200gat
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    r, c, s = b200gat.build_ii_knn(torch.from_numpy(g["embeddings"]).cuda(), k=int(g["k"]), min_similarity=float(g["min_similarity"]))
    r, c, s = r.cpu().numpy(), c.cpu().numpy(), s.cpu().numpy()
    gs = g["sims"]
    np.testing.assert_array_equal(r, g["rows"])                                     # same number of edges in every row
    np.testing.assert_allclose(s, gs, rtol=0, atol=3e-6)                            # same similarity at every position
    # same neighbour at every position, except inside runs of similarities closer than 1e-5 (order may flip there) ...
    near = np.zeros(len(gs), dtype=bool)
    close = (np.abs(np.diff(gs)) < 1e-5) & (np.diff(g["rows"]) == 0)
    near[:-1] |= close
    near[1:] |= close
    assert np.all((c == g["cols"]) | near)
    assert near.mean() < 0.05
    # ... and the neighbour SETS are identical (the fixtures keep the k-th / (k+1)-th and the min_similarity cut > 1e-5 apart)
    key = lambda rr, cc: np.sort(rr.astype(np.int64) * (1 << 32) + cc)
    np.testing.assert_array_equal(key(r, c), key(g["rows"], g["cols"]))


def _assert_exact_topk(emb, r, c, s, k):
    n = emb.shape[0]
    en = emb.astype(np.float64)
    en /= np.linalg.norm(en, axis=1, keepdims=True)
    full = en @ en.T
    np.fill_diagonal(full, -np.inf)
    assert len(r) == n * k
    np.testing.assert_allclose(s, full[r, c], rtol=0, atol=3e-6)
    kth = -np.sort(-full, axis=1)[:, k - 1]
    assert np.all(full[r, c] >= kth[r] - 5e-6)            # every emitted neighbour is among the row's k best (up to rounding ties)


@pytest.mark.parametrize("mode", ["append", "lists"])
def test_knn_survives_index_sorted_clusters(mode, monkeypatch):
    """Adversarial layout for sampled admission thresholds: items are stored cluster by cluster, one cluster per 128-row
    column block, so the sampled blocks (every 16th / every 8th) show each row a threshold taken from OTHER clusters only or --
    for rows of a sampled block -- from its own cluster only.  The result must still be the exact top-k, in the append
    pipeline (default at this size) and with the register-list kernel forced."""
    import b200gat
    if mode == "lists":
        monkeypatch.setenv("B200GAT_KNN_MODE", "lists")
    rng = np.random.default_rng(5)
    n_clusters, per = 72, 128
    centers = rng.standard_normal((n_clusters, 128))
    emb = (np.repeat(centers, per, axis=0) + 0.25 * rng.standard_normal((n_clusters * per, 128))).astype(np.float32)
    k = 20
    r, c, s = b200gat.build_ii_knn(torch.from_numpy(emb).cuda(), k=k, min_similarity=-1.0)
    _assert_exact_topk(emb, r.cpu().numpy(), c.cpu().numpy(), s.cpu().numpy(), k)


@pytest.mark.parametrize("max_overflow", [None, 100000])
def test_knn_append_lists_overflow(max_overflow, monkeypatch):
    """A clump of 1,000 near-duplicates that no sampled column block shows (it sits in blocks 1..7 and 9): each of its rows
    appends more entries than its list holds in the full sweep.  Default: more than 64 such rows -> the register-list kernel behind the
    device-side gate redoes the sweep; with the limit raised the overflowed rows are handed to the exact path instead.  Either way
    the result is the exact top-k."""
    import b200gat
    if max_overflow is not None:
        monkeypatch.setenv("B200GAT_KNN_MAX_OVERFLOW", str(max_overflow))
    rng = np.random.default_rng(11)
    n, d, k = 9000, 128, 20
    emb = rng.standard_normal((n, d)).astype(np.float32)
    clump = np.r_[128:1024, 1152:1256]                    # 896 + 104 rows, none of them in a block with index % 8 == 0
    emb[clump] = (rng.standard_normal(d) + 0.3 * rng.standard_normal((len(clump), d))).astype(np.float32)
    stats = {}
    idx, sim, counts = b200gat.knn_neighbors(torch.from_numpy(emb).cuda(), k, -1.0, stats=stats)
    r, c, s = b200gat.build_ii_knn(torch.from_numpy(emb).cuda(), k=k, min_similarity=-1.0)
    _assert_exact_topk(emb, r.cpu().numpy(), c.cpu().numpy(), s.cpu().numpy(), k)
    if max_overflow is not None:
        assert int(stats["exact_rows"].item()) >= len(clump)      # the overflowed rows went through the exact path
