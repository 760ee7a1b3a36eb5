"""Graph structure (CSR by destination + CSC by source) built on the device, and its cache.

The reference passes the same ``edge_index`` tensor object to every layer of every epoch
(scripts/train_gat_custom.py:330,349), so the structure is built once per tensor and cached.
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass
from typing import Dict, Tuple

import torch

from . import _lib


@dataclass
class GraphStructure:
    n_nodes: int
    n_edges: int
    rowptr: torch.Tensor    # int32 [N+1]  CSR by destination
    col: torch.Tensor       # int32 [E]    source of each edge, CSR order
    perm: torch.Tensor      # int32 [E]    original edge id, CSR order  (== dst.argsort(stable=True))
    colptr: torch.Tensor    # int32 [N+1]  CSC by source
    row: torch.Tensor       # int32 [E]    destination of each edge, CSC order
    perm_csc: torch.Tensor  # int32 [E]
    csr2csc: torch.Tensor   # int32 [E]    CSC position of the edge at each CSR position
    sched_fwd: "_lib.Schedule" = None   # destination rows by descending in-degree, long rows split into segments
    sched_bwd: "_lib.Schedule" = None   # source rows by descending out-degree


def build_graph(edge_index: torch.Tensor, n_nodes: int, check: bool = True) -> GraphStructure:
    """Device COO -> CSR/CSC (b200gat_build_graph).  ``edge_index``: int64 [2, E] on a CUDA device, row 0 = source,
    row 1 = destination (scripts/train_gat_custom.py:76,78)."""
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise RuntimeError(f"edge_index must be [2, E], got {tuple(edge_index.shape)}")
    if edge_index.dtype != torch.int64:
        raise RuntimeError(f"edge_index must be int64 (as built by build_edge_index), got {edge_index.dtype}")
    if not edge_index.is_cuda:
        raise RuntimeError("edge_index must live on a CUDA device (there is no CPU path)")
    ei = edge_index.contiguous()
    dev = ei.device
    n_edges = int(ei.shape[1])
    i32 = dict(dtype=torch.int32, device=dev)
    g = GraphStructure(n_nodes, n_edges,
                       torch.empty(n_nodes + 1, **i32), torch.empty(n_edges, **i32), torch.empty(n_edges, **i32),
                       torch.empty(n_nodes + 1, **i32), torch.empty(n_edges, **i32), torch.empty(n_edges, **i32),
                       torch.empty(n_edges, **i32))
    ws_bytes = _lib.graph_workspace_bytes(n_nodes, n_edges)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    n_bad = torch.empty(1, **i32)
    with torch.cuda.device(dev):
        _lib.call("b200gat_build_graph", _lib.ptr(ei), n_edges, n_nodes, _lib.ptr(g.rowptr), _lib.ptr(g.col),
                  _lib.ptr(g.perm), _lib.ptr(g.colptr), _lib.ptr(g.row), _lib.ptr(g.perm_csc), _lib.ptr(g.csr2csc),
                  _lib.ptr(n_bad), _lib.ptr(ws), ws_bytes, _lib.stream())
    g.sched_fwd = _lib.make_schedule(g.rowptr, 0, n_nodes, n_edges + 1)
    g.sched_bwd = _lib.make_schedule(g.colptr, 0, n_nodes, n_edges + 1)
    if check:
        bad = int(n_bad.item())
        if bad:
            raise IndexError(f"edge_index has {bad} edge(s) with an endpoint outside [0, {n_nodes})")
    return g


_cache: Dict[int, Tuple[weakref.ref, tuple, GraphStructure]] = {}


def graph_for(edge_index: torch.Tensor, n_nodes: int) -> GraphStructure:
    """Cached :func:`build_graph`, keyed on the identity of the ``edge_index`` tensor object (plus its version
    counter, storage pointer and shape, so in-place edits and re-allocations rebuild)."""
    key = id(edge_index)
    sig = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), n_nodes, str(edge_index.device))
    hit = _cache.get(key)
    if hit is not None and hit[0]() is edge_index and hit[1] == sig:
        return hit[2]
    g = build_graph(edge_index, n_nodes)
    _cache[key] = (weakref.ref(edge_index, lambda _r, k=key: _cache.pop(k, None)), sig, g)
    return g


def clear_graph_cache() -> None:
    _cache.clear()


def build_edge_index(n_users: int, n_items: int, train_pos_idx) -> torch.Tensor:
    """Same contract as the reference's ``build_edge_index`` (scripts/train_gat_custom.py:166-175): interleaved
    ``(u -> n_users+i), (n_users+i -> u)`` per interaction in dict/array order, int64 [2, E] on the host.
    Vectorised per user instead of the reference's per-edge Python loop."""
    import numpy as np
    users = list(train_pos_idx.keys())
    lens = np.fromiter((len(train_pos_idx[u]) for u in users), dtype=np.int64, count=len(users))
    if lens.sum() == 0:
        return torch.zeros((2, 0), dtype=torch.long)
    items = np.concatenate([np.asarray(train_pos_idx[u], dtype=np.int64) for u in users]) + n_users
    uu = np.repeat(np.asarray(users, dtype=np.int64), lens)
    out = np.empty((2, 2 * items.size), dtype=np.int64)
    out[0, 0::2], out[1, 0::2] = uu, items
    out[0, 1::2], out[1, 1::2] = items, uu
    return torch.from_numpy(out)


def union_edge_index(ui_edge_index: torch.Tensor, n_users: int, ii_rows: torch.Tensor, ii_cols: torch.Tensor) -> torch.Tensor:
    """The graph the hot path runs on (BASELINE.json north_star): the User-Item block of ``build_edge_index`` followed by the
    directed Item-Item kNN block.  ``ii_rows`` / ``ii_cols`` are the COO of graphs/build_ii_knn.py:103-111 (row = item,
    col = neighbour, item ids), i.e. what ``b200gat.build_ii_knn`` returns or ``scipy.sparse.load_npz(ii_edges_*.npz)``
    holds; the edge for a COO entry is ``n_users + row -> n_users + col``, appended after the U-I block in COO order (the
    reference never defines this union; SURVEY.md section 8d records the decision).  int64 [2, E] on ``ui_edge_index``'s
    device; duplicates are kept, nothing is sorted."""
    if ui_edge_index.dim() != 2 or ui_edge_index.shape[0] != 2:
        raise ValueError(f"ui_edge_index must be [2, E], got {tuple(ui_edge_index.shape)}")
    if ii_rows.shape != ii_cols.shape or ii_rows.dim() != 1:
        raise ValueError("ii_rows and ii_cols must be 1-D tensors of equal length")
    dev = ui_edge_index.device
    ii = torch.stack([ii_rows.to(device=dev, dtype=torch.int64), ii_cols.to(device=dev, dtype=torch.int64)]) + n_users
    return torch.cat([ui_edge_index.to(torch.int64), ii], dim=1).contiguous()
