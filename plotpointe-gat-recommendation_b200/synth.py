"""Synthetic graphs of the reference's shape: User-Item interactions in the reference's interleaved
edge order (scripts/train_gat_custom.py:166-175) followed by a directed Item-Item k-NN block
(graphs/build_ii_knn.py:103-111: row = item, col = neighbour)."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

CONFIGS = {
    # name: (n_users, n_items, n_interactions, k)
    "cfg1": (10_000, 20_000, 200_000, 20),
    "amazon": (192_403, 498_196, 1_689_116, 20),
    "cfg5": (10_000_000, 20_000_000, 200_000_000, 20),      # BASELINE config 5: 30 M nodes, 800 M edges
    "tiny": (300, 500, 4_000, 8),
}


def make_graph(n_users: int, n_items: int, n_inter: int, k: int = 20, seed: int = 42, item_zipf: float = 0.9,
               item_shift: float = 40.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (edge_index int64 [2, 2*n_inter + k*n_items] on the host, item_feats fp32 [n_items, 128]).

    Users get one interaction each plus a geometric-ish remainder (mean n_inter/n_users); items are drawn from a
    shifted power law so item in-degree is heavy-tailed (std >> mean, as the reference's graph stats report);
    duplicate (u, i) pairs are kept.  Each item gets k distinct random neighbours (exact cosine kNN is O(n_items^2)
    and only used for small graphs, see oracle.build_ii_knn)."""
    rng = np.random.default_rng(seed)
    extra = n_inter - n_users
    assert extra >= 0
    w = rng.exponential(1.0, size=n_users)
    deg = 1 + rng.multinomial(extra, w / w.sum())
    users = np.repeat(np.arange(n_users, dtype=np.int64), deg)
    ranks = np.arange(1, n_items + 1, dtype=np.float64)
    pop = 1.0 / (ranks + item_shift) ** item_zipf
    pop /= pop.sum()
    item_of_rank = rng.permutation(n_items)
    items = item_of_rank[rng.choice(n_items, size=n_inter, p=pop)].astype(np.int64) + n_users
    ui = np.empty((2, 2 * n_inter), dtype=np.int64)
    ui[0, 0::2], ui[1, 0::2] = users, items
    ui[0, 1::2], ui[1, 1::2] = items, users
    # k distinct neighbours != self: offsets in [1, n_items) without replacement per row (vectorised via argpartition
    # of random keys would be O(n_items^2); draw k offsets and de-duplicate by sorting + bumping instead)
    off = np.sort(rng.integers(1, n_items - k + 1, size=(n_items, k)), axis=1) + np.arange(k)[None, :]
    rows = np.repeat(np.arange(n_items, dtype=np.int64), k)
    cols = (rows + off.reshape(-1)) % n_items
    ii = np.stack([rows + n_users, cols + n_users])
    edge_index = torch.from_numpy(np.concatenate([ui, ii], axis=1))
    feats = rng.standard_normal((n_items, 128)).astype(np.float32)
    feats /= np.linalg.norm(feats, axis=1, keepdims=True)
    return edge_index, torch.from_numpy(feats)


def make_triples(n_users: int, n_items: int, n_triples: int, seed: int = 42):
    rng = np.random.default_rng(seed + 1)
    return (torch.from_numpy(rng.integers(0, n_users, size=n_triples)),
            torch.from_numpy(rng.integers(0, n_items, size=n_triples)),
            torch.from_numpy(rng.integers(0, n_items, size=n_triples)))


def make_graph_device(n_users: int, n_items: int, n_inter: int, k: int, device, seed: int = 42, item_zipf: float = 0.9,
                      item_shift: float = 40.0):
    """The same construction as :func:`make_graph`, drawn with torch on ``device`` (a different random stream, the same
    distributions, edge order and shapes).  For BASELINE config 5 at full size (30 M nodes, 800 M edges) the host generator
    needs ~5 minutes and ~77 GB per process; this one takes seconds.  Returns (edge_index int64 [2, E], item_feats fp32
    [n_items, 128]) on ``device``.  NOT bitwise reproducible between processes at large sizes (torch.cumsum on CUDA is not
    deterministic, which moves a few inverse-CDF draws): a multi-GPU job draws the edge list on rank 0 and broadcasts it."""
    g = torch.Generator(device=device).manual_seed(seed)
    f64 = dict(dtype=torch.float64, device=device)
    extra = n_inter - n_users
    assert extra >= 0
    w = -torch.log1p(-torch.rand(n_users, generator=g, **f64))                     # Exp(1) weights, as the host generator
    cdf = torch.cumsum(w / w.sum(), 0)
    users = torch.cat([torch.arange(n_users, device=device),
                       torch.searchsorted(cdf, torch.rand(extra, generator=g, **f64)).clamp_(max=n_users - 1)])
    users = torch.sort(users).values                                             # edges grouped by user (build_edge_index order)
    del w, cdf
    ranks = torch.arange(1, n_items + 1, **f64)
    pop = 1.0 / (ranks + item_shift) ** item_zipf
    cdf = torch.cumsum(pop / pop.sum(), 0)
    del ranks, pop
    item_of_rank = torch.randperm(n_items, generator=g, device=device)
    items = torch.empty(n_inter, dtype=torch.int64, device=device)
    step = 1 << 26
    for lo in range(0, n_inter, step):                                            # chunked: the fp64 uniforms are transient
        n = min(step, n_inter - lo)
        items[lo:lo + n] = item_of_rank[torch.searchsorted(cdf, torch.rand(n, generator=g, **f64)).clamp_(max=n_items - 1)]
    del cdf, item_of_rank
    items += n_users
    e_ui, e_ii = 2 * n_inter, k * n_items
    ei = torch.empty((2, e_ui + e_ii), dtype=torch.int64, device=device)
    ei[0, 0:e_ui:2], ei[1, 0:e_ui:2] = users, items
    ei[0, 1:e_ui:2], ei[1, 1:e_ui:2] = items, users
    del users, items
    step_rows = max((1 << 25) // k, 1)
    ar_k = torch.arange(k, device=device)
    for lo in range(0, n_items, step_rows):
        n = min(step_rows, n_items - lo)
        off = torch.sort(torch.randint(1, n_items - k + 1, (n, k), generator=g, device=device), dim=1).values + ar_k
        rows = torch.arange(lo, lo + n, device=device).unsqueeze(1).expand(n, k)
        ei[0, e_ui + lo * k:e_ui + (lo + n) * k] = (rows + n_users).reshape(-1)
        ei[1, e_ui + lo * k:e_ui + (lo + n) * k] = ((rows + off) % n_items + n_users).reshape(-1)
    return ei, make_feats_device(n_items, device, seed)


def make_feats_device(n_items: int, device, seed: int = 42) -> torch.Tensor:
    """Unit-norm random item features [n_items, 128] drawn on ``device`` (element-wise Philox draws and a row norm: the same
    values on every device of the same type, so each rank of a multi-GPU run can make its own copy)."""
    g = torch.Generator(device=device).manual_seed(seed + 7)
    feats = torch.randn((n_items, 128), generator=g, device=device)
    feats /= feats.norm(dim=1, keepdim=True)
    return feats
