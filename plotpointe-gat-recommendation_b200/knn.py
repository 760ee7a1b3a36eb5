"""Item-Item cosine kNN on the GPU -- the graph source of graphs/build_ii_knn.py:54-111."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib


def knn_neighbors(embeddings: torch.Tensor, k: int = 20, min_similarity: float = 0.3, stats: Optional[dict] = None):
    """Top-k cosine neighbours of every row.  Returns (nbr_idx int32 [n, k], nbr_sim fp32 [n, k], counts int32 [n]):
    rows are in descending similarity, self excluded; ``counts`` = how many entries pass ``>= min_similarity``.
    ``embeddings`` is [n, 128] (fused features) or [n, 384] (text embeddings), fp32, on the GPU.  If ``stats`` is a dict,
    ``stats["exact_rows"]`` receives a device int32 tensor: how many rows the bf16 candidate pass could not prove exact
    and were therefore recomputed against all columns (dense clumps of near-duplicates); the result is exact either way."""
    if not embeddings.is_cuda:
        raise RuntimeError("b200gat knn: embeddings must be a CUDA tensor (there is no CPU fallback)")
    emb = _lib._f32(embeddings, "embeddings").contiguous()
    n, d = emb.shape
    out = ctypes.c_size_t(0)
    _lib._check(_lib._lib.b200gat_knn_workspace_bytes(n, d, ctypes.byref(out)), "knn_workspace_bytes")
    ws = torch.empty(out.value, dtype=torch.uint8, device=emb.device)
    idx = torch.empty((n, k), dtype=torch.int32, device=emb.device)
    sim = torch.empty((n, k), dtype=torch.float32, device=emb.device)
    counts = torch.empty(n, dtype=torch.int32, device=emb.device)
    unsafe = torch.empty(1, dtype=torch.int32, device=emb.device)
    with torch.cuda.device(emb.device):
        _lib.call("b200gat_knn_cosine_f32", _lib.ptr(emb), n, d, k, float(min_similarity), _lib.ptr(idx), _lib.ptr(sim), _lib.ptr(counts),
                  _lib.ptr(unsafe), _lib.ptr(ws), out.value, _lib.stream())
    if stats is not None:
        stats["exact_rows"] = unsafe
    return idx, sim, counts


def build_ii_knn(embeddings: torch.Tensor, k: int = 20, min_similarity: float = 0.3) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Same output as the reference builder (graphs/build_ii_knn.py:103-111): COO ``(rows int32, cols int32, sims fp32)``,
    row = item, col = neighbour, items in order, neighbours in descending similarity, weak links (< min_similarity) dropped."""
    idx, sim, counts = knn_neighbors(embeddings, k, min_similarity)
    keep = torch.arange(k, device=idx.device).unsqueeze(0) < counts.unsqueeze(1)
    rows = torch.arange(idx.shape[0], dtype=torch.int32, device=idx.device).unsqueeze(1).expand(-1, k)
    return rows[keep], idx[keep], sim[keep]
