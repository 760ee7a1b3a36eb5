"""CPU oracle for the GAT hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file restates, in plain torch/numpy CPU ops, the arithmetic of the reference's GAT
training hot path so the CUDA kernels can be checked against it. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it. The product package (``plotpointe-gat-recommendation_b200``) never does.

Pinning status
--------------
* custom dialect (``simple_gat_layer`` / ``custom_gat_forward`` / ``bpr_loss`` / ``bce_loss`` /
  ``build_edge_index``): PINNED. ``oracle/make_golden.py`` imports the reference's unmodified
  ``scripts/train_gat_custom.py`` in the build container, runs it on seeded inputs and stores
  inputs + outputs + gradients under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
  this file against those vectors.
* PyG dialect (``gatconv``): the arithmetic lives in the third-party ``torch-geometric``
  package, which the reference installs un-pinned (docker/Dockerfile:21-23) and which is absent
  from /root/reference and from this image.  It is restated from the published semantics of
  ``torch_geometric.nn.GATConv`` (2.6.x at the date of the reference's runs) for the one call the
  reference makes (scripts/train_gat_pyg.py:77,87).  PARITY UNPINNED for this dialect (no reference-held vector can
  exist here).  Two independent anchors stand in for one: (1) the cross-check against the pinned custom dialect on inputs
  where both coincide (``tests/test_oracle_golden.py::test_gatconv_matches_custom_when_unclamped``); (2)
  ``gatconv_dense``, the textbook dense-adjacency masked-softmax statement of the same layer, checked against ``gatconv``
  in fp64 (output and every gradient) for heads 1/2/4, non-zero bias, parallel edges, rows without in-edges and logits far
  outside [-10, 10] (``test_gatconv_matches_dense_formulation``).
* ``eval_ranks`` / ``ranking_metrics`` / ``sample_eval_candidates`` (next row f2): PINNED.  ``make_golden.py`` runs
  the reference's own ``eval_sampled`` (scripts/train_gat_custom.py:184-210) under a fixed numpy seed and stores its
  metrics; the test replays the same numpy stream through ``sample_eval_candidates`` and must reproduce them.
* ``build_ii_knn`` (next row f1): PINNED.  graphs/build_ii_knn.py:56-99 only exists inside a ``main()`` that talks to
  GCS, so ``make_golden.py knn`` runs that file AS A SCRIPT (runpy) with a stand-in storage client that hands over a
  local .npy, and stores the script's own output (``tests/golden/knn_128.npz``, ``knn_384.npz``); the restatement must
  reproduce the edge list (rows, neighbours, order) exactly and the similarities to fp32 rounding.

All functions are dtype-generic (float32 or float64) and differentiable through torch autograd,
which is exactly how the reference obtains its gradients (``loss.backward()``,
scripts/train_gat_custom.py:361).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# a1: edge list + CSR ordering
# --------------------------------------------------------------------------------------


def build_edge_index(n_users: int, n_items: int, train_pos_idx: Dict[int, Sequence[int]]) -> torch.Tensor:
    """Interleaved symmetric U<->I COO list, follows scripts/train_gat_custom.py:166-175
    (identical in scripts/train_gat_pyg.py:139-147).

    For every user (dict order) and every item of that user (array order) two edges are emitted,
    ``(u -> n_users+i)`` then ``(n_users+i -> u)``.  No dedup, no sort.  Row 0 = source, row 1 =
    destination (train_gat_custom.py:78).
    """
    src_parts, dst_parts = [], []
    for u, items in train_pos_idx.items():
        it = np.asarray(items, dtype=np.int64) + n_users
        uu = np.full(it.shape, int(u), dtype=np.int64)
        s = np.empty(2 * it.size, dtype=np.int64)
        d = np.empty(2 * it.size, dtype=np.int64)
        s[0::2], d[0::2] = uu, it
        s[1::2], d[1::2] = it, uu
        src_parts.append(s)
        dst_parts.append(d)
    if not src_parts:
        return torch.zeros((2, 0), dtype=torch.long)
    return torch.from_numpy(np.stack([np.concatenate(src_parts), np.concatenate(dst_parts)]))


def csr_by_dst(edge_index: torch.Tensor, n_nodes: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """CSR oracle: stable sort of the COO list by destination.

    Returns (rowptr[N+1], col[E] = src in sorted order, perm[E] = original edge ids).  Duplicates
    are kept and the intra-row order is the original edge order -- the order in which the
    reference's sequential CPU ``index_add_`` (train_gat_custom.py:92) visits the edges of a row.
    """
    src = edge_index[0].numpy()
    dst = edge_index[1].numpy()
    perm = np.argsort(dst, kind="stable")
    rowptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(dst, minlength=n_nodes), out=rowptr[1:])
    return rowptr, src[perm], perm


def csc_by_src(edge_index: torch.Tensor, n_nodes: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Twin of :func:`csr_by_dst` keyed on the source row (used by the backward pass)."""
    src = edge_index[0].numpy()
    dst = edge_index[1].numpy()
    perm = np.argsort(src, kind="stable")
    colptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n_nodes), out=colptr[1:])
    return colptr, dst[perm], perm


# --------------------------------------------------------------------------------------
# a4-a9: the custom layer
# --------------------------------------------------------------------------------------


def simple_gat_layer(x: torch.Tensor, edge_index: torch.Tensor, lin_weight: torch.Tensor,
                     a_src: torch.Tensor, a_dst: torch.Tensor,
                     keep_mask: Optional[torch.Tensor] = None, p_drop: float = 0.0) -> torch.Tensor:
    """Restates SimpleGATLayer.forward, scripts/train_gat_custom.py:75-93.

    ``keep_mask`` ([E] of 0/1) stands in for ``torch.nn.Dropout`` (:89): alpha is multiplied by
    ``keep/(1-p)``.  ``None`` is eval mode.
    """
    h = x @ lin_weight.t()                                              # :77 (bias-free Linear)
    j, i = edge_index[0], edge_index[1]                                 # :78
    logit = (h.index_select(0, j) * a_src).sum(-1) + (h.index_select(0, i) * a_dst).sum(-1)  # :79
    logit = torch.nn.functional.leaky_relu(logit, 0.2)                  # :80
    logit = logit.clamp(-10.0, 10.0)                                    # :82
    num = logit.exp()                                                   # :83
    den = torch.zeros(h.shape[0], dtype=num.dtype).scatter_add_(0, i, num)   # :85-87
    alpha = num / (den.index_select(0, i) + 1e-9)                       # :88
    if keep_mask is not None:
        alpha = alpha * keep_mask.to(alpha.dtype) / (1.0 - p_drop)      # :89
    out = torch.zeros_like(h)                                           # :91
    out.index_add_(0, i, alpha.unsqueeze(-1) * h.index_select(0, j))    # :92
    return out


def node_features(user_emb: torch.Tensor, item_proj_w: torch.Tensor, item_proj_b: torch.Tensor,
                  item_feats: torch.Tensor) -> torch.Tensor:
    """CustomGAT.node_features, train_gat_custom.py:105-109 (same in train_gat_pyg.py:79-82)."""
    return torch.cat([user_emb, item_feats @ item_proj_w.t() + item_proj_b], dim=0)


def custom_gat_forward(state: Dict[str, torch.Tensor], item_feats: torch.Tensor,
                       edge_index: torch.Tensor, p_drop: float = 0.0,
                       generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """CustomGAT.forward (train_gat_custom.py:111-115) over a reference-named state dict.  ``p_drop > 0`` is train
    mode: a fresh keep mask per layer, as ``self.drop(alpha)`` (:89) draws one."""
    x = node_features(state["user_emb.weight"], state["item_proj.weight"], state["item_proj.bias"], item_feats)
    layer = 0
    while f"layers.{layer}.lin.weight" in state:
        keep = None
        if p_drop > 0.0:
            keep = torch.rand(edge_index.shape[1], generator=generator) >= p_drop
        x = simple_gat_layer(x, edge_index, state[f"layers.{layer}.lin.weight"],
                             state[f"layers.{layer}.a_src"], state[f"layers.{layer}.a_dst"], keep_mask=keep, p_drop=p_drop)
        layer += 1
    return x


# --------------------------------------------------------------------------------------
# a10: PyG GATConv(heads=H, concat=False, add_self_loops=False)
# --------------------------------------------------------------------------------------


def gatconv(x: torch.Tensor, edge_index: torch.Tensor, lin_weight: torch.Tensor, att_src: torch.Tensor,
            att_dst: torch.Tensor, bias: Optional[torch.Tensor], heads: int,
            negative_slope: float = 0.2, keep_mask: Optional[torch.Tensor] = None,
            p_drop: float = 0.0) -> torch.Tensor:
    """Restates ``GATConv(C_in, C, heads=H, concat=False, add_self_loops=False)(x, edge_index)``
    as called at scripts/train_gat_pyg.py:77,87.

    lin_weight [H*C, F_in], att_* [1, H, C], bias [C].  Softmax per destination with the running
    maximum subtracted (detached) and 1e-16 added to the denominator; heads averaged; bias added.
    ``keep_mask`` is [E, H].
    """
    n = x.shape[0]
    h = (x @ lin_weight.t()).view(n, heads, -1)
    s_src = (h * att_src).sum(-1)                                        # [N, H]
    s_dst = (h * att_dst).sum(-1)
    j, i = edge_index[0], edge_index[1]
    z = torch.nn.functional.leaky_relu(s_src.index_select(0, j) + s_dst.index_select(0, i), negative_slope)
    zmax = torch.full((n, heads), float("-inf"), dtype=z.dtype)
    zmax = zmax.scatter_reduce(0, i.unsqueeze(-1).expand(-1, heads), z.detach(), reduce="amax", include_self=True)
    zmax = torch.where(torch.isfinite(zmax), zmax, torch.zeros_like(zmax))
    num = (z - zmax.index_select(0, i)).exp()
    den = torch.zeros((n, heads), dtype=z.dtype).index_add_(0, i, num)
    alpha = num / (den.index_select(0, i) + 1e-16)
    if keep_mask is not None:
        alpha = alpha * keep_mask.to(alpha.dtype) / (1.0 - p_drop)
    msg = alpha.unsqueeze(-1) * h.index_select(0, j)                     # [E, H, C]
    out = torch.zeros_like(h).index_add_(0, i, msg).mean(dim=1)          # concat=False
    if bias is not None:
        out = out + bias
    return out


def gatconv_dense(x: torch.Tensor, edge_index: torch.Tensor, lin_weight: torch.Tensor, att_src: torch.Tensor,
                  att_dst: torch.Tensor, bias: Optional[torch.Tensor], heads: int,
                  negative_slope: float = 0.2) -> torch.Tensor:
    """Second, structurally independent statement of the same ``GATConv`` call (eval mode): dense adjacency and a masked
    softmax over an [N, N, H] logit tensor instead of gathers and scatters -- the textbook GAT formula
    ``alpha_ij = softmax_j(LeakyReLU(a_src.h_j + a_dst.h_i))`` over the in-neighbours j of i (Velickovic et al. 2018, eq. 3,
    which is what ``GATConv`` documents).  Parallel edges count separately in PyG's edge softmax, so the multiplicity
    M[i, j] of edge j->i enters as ``+ log M`` inside the softmax.  O(N^2 H) memory: small graphs only.  No eps in the
    denominator: after the max-subtraction PyG's denominator is >= 1, so its ``+1e-16`` is below fp64 resolution.
    Used only to cross-check :func:`gatconv` (tests/test_oracle_golden.py)."""
    n = x.shape[0]
    h = (x @ lin_weight.t()).view(n, heads, -1)                          # [N, H, C]
    s_src = torch.einsum("nhc,hc->nh", h, att_src.view(heads, -1))
    s_dst = torch.einsum("nhc,hc->nh", h, att_dst.view(heads, -1))
    mult = torch.zeros((n, n), dtype=x.dtype)
    mult.index_put_((edge_index[1], edge_index[0]), torch.ones(edge_index.shape[1], dtype=x.dtype), accumulate=True)
    z = torch.nn.functional.leaky_relu(s_dst.unsqueeze(1) + s_src.unsqueeze(0), negative_slope)    # [i, j, H]
    z = z + torch.log(mult).unsqueeze(-1)                                # -inf where there is no edge j -> i
    has_in = (mult.sum(1) > 0).view(n, 1, 1)
    alpha = torch.softmax(torch.where(has_in, z, torch.zeros_like(z)), dim=1)
    alpha = torch.where(has_in, alpha, torch.zeros_like(alpha))          # rows without in-edges aggregate nothing
    out = torch.einsum("ijh,jhc->ihc", alpha, h).mean(dim=1)
    if bias is not None:
        out = out + bias
    return out


def pyg_gat_forward(state: Dict[str, torch.Tensor], item_feats: torch.Tensor, edge_index: torch.Tensor,
                    heads: int, p_drop: float = 0.0, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """PyGGAT.forward (train_gat_pyg.py:84-88) over a PyG(>=2.5)-named state dict.  ``p_drop > 0`` is train mode:
    a fresh Bernoulli keep mask per layer, as ``F.dropout(alpha, p, training=True)`` inside GATConv draws one."""
    x = node_features(state["user_emb.weight"], state["item_proj.weight"], state["item_proj.bias"], item_feats)
    layer = 0
    while f"convs.{layer}.lin.weight" in state:
        p = f"convs.{layer}."
        keep = None
        if p_drop > 0.0:
            keep = torch.rand((edge_index.shape[1], heads), generator=generator) >= p_drop
        x = gatconv(x, edge_index, state[p + "lin.weight"], state[p + "att_src"], state[p + "att_dst"],
                    state[p + "bias"], heads, keep_mask=keep, p_drop=p_drop)
        layer += 1
    return x


def init_custom_state(n_users: int, n_items: int, item_feat_dim: int, hidden: int, layers: int) -> Dict[str, torch.Tensor]:
    """Parameters of the reference ``CustomGAT`` drawn in its constructor's order with its initialisers
    (scripts/train_gat_custom.py:64-71,97-103), under the caller's ``torch.manual_seed``: a reference-named state dict
    without instantiating any module (bench.py's CPU arm must not import the product package)."""
    st = {"user_emb.weight": torch.nn.Embedding(n_users, hidden).weight.detach()}
    torch.nn.init.normal_(st["user_emb.weight"], std=0.1)
    proj = torch.nn.Linear(item_feat_dim, hidden)
    st["item_proj.weight"], st["item_proj.bias"] = proj.weight.detach(), proj.bias.detach()
    for l in range(layers):
        w = torch.nn.Linear(hidden, hidden, bias=False).weight.detach()
        a_s, a_d = torch.empty(hidden), torch.empty(hidden)
        torch.nn.init.xavier_uniform_(w)
        torch.nn.init.xavier_uniform_(a_s.unsqueeze(0))
        torch.nn.init.xavier_uniform_(a_d.unsqueeze(0))
        st[f"layers.{l}.lin.weight"], st[f"layers.{l}.a_src"], st[f"layers.{l}.a_dst"] = w, a_s, a_d
    return st


def init_pyg_state(n_users: int, n_items: int, item_feat_dim: int, hidden: int, layers: int, heads: int) -> Dict[str, torch.Tensor]:
    """Same for ``PyGGAT`` (scripts/train_gat_pyg.py:68-77) with GATConv's published initialisers: glorot for
    ``lin.weight`` / ``att_src`` / ``att_dst``, zeros for ``bias``."""
    import math
    st = {"user_emb.weight": torch.nn.Embedding(n_users, hidden).weight.detach()}
    torch.nn.init.normal_(st["user_emb.weight"], std=0.1)
    proj = torch.nn.Linear(item_feat_dim, hidden)
    st["item_proj.weight"], st["item_proj.bias"] = proj.weight.detach(), proj.bias.detach()

    def glorot(t):
        bound = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
        return t.uniform_(-bound, bound)
    for l in range(layers):
        w = torch.nn.Linear(hidden, heads * hidden, bias=False).weight.detach()
        st[f"convs.{l}.att_src"] = torch.empty(1, heads, hidden)
        st[f"convs.{l}.att_dst"] = torch.empty(1, heads, hidden)
        st[f"convs.{l}.bias"] = torch.zeros(hidden)
        st[f"convs.{l}.lin.weight"] = glorot(w)
        glorot(st[f"convs.{l}.att_src"])
        glorot(st[f"convs.{l}.att_dst"])
    return st


# --------------------------------------------------------------------------------------
# a11-a13: ranking losses
# --------------------------------------------------------------------------------------


def pos_neg_scores(z: torch.Tensor, n_users: int, u: torch.Tensor, i: torch.Tensor, j: torch.Tensor):
    """train_gat_custom.py:350-353."""
    users, items = z[:n_users], z[n_users:]
    uu = users.index_select(0, u)
    return (uu * items.index_select(0, i)).sum(-1), (uu * items.index_select(0, j)).sum(-1)


def bpr_loss(z, n_users, u, i, j):
    """train_gat_custom.py:354-355: mean(-log(sigmoid(pos-neg) + 1e-8))."""
    pos, neg = pos_neg_scores(z, n_users, u, i, j)
    return -(torch.sigmoid(pos - neg) + 1e-8).log().mean()


def bce_loss(z, n_users, u, i, j):
    """train_gat_custom.py:356-359: BCE-with-logits over cat[pos, neg] vs cat[1, 0]."""
    pos, neg = pos_neg_scores(z, n_users, u, i, j)
    logits = torch.cat([pos, neg])
    labels = torch.cat([torch.ones_like(pos), torch.zeros_like(neg)])
    return torch.nn.functional.binary_cross_entropy_with_logits(logits, labels)


# --------------------------------------------------------------------------------------
# f1 (next row): item-item cosine kNN
# --------------------------------------------------------------------------------------


def build_ii_knn(embeddings: np.ndarray, k: int = 20, min_similarity: float = 0.3, batch_size: int = 1000):
    """graphs/build_ii_knn.py:56-111: cosine top-k per item, self excluded, descending similarity,
    rows filtered by ``>= min_similarity``; COO (row=item, col=neighbour, sim) int32/int32/float32.

    The reference calls sklearn's ``cosine_similarity`` on already-normalised rows, which
    normalises them a second time before the dot product; restated here in numpy.
    """
    e = embeddings / (np.linalg.norm(embeddings, axis=1, keepdims=True) + 1e-8)          # :56-57

    def _sk_normalize(a):
        nrm = np.sqrt((a * a).sum(axis=1, keepdims=True))
        nrm[nrm == 0.0] = 1.0
        return a / nrm

    en = _sk_normalize(e)
    rows, cols, sims = [], [], []
    n = e.shape[0]
    for start in range(0, n, batch_size):
        stop = min(start + batch_size, n)
        sim = en[start:stop] @ en.T                                                      # :76
        for r in range(stop - start):
            s = sim[r]
            s[start + r] = -np.inf                                                       # :83
            top = np.argpartition(s, -k)[-k:]                                            # :86
            top = top[np.argsort(s[top])[::-1]]                                          # :87
            ts = s[top]
            keep = ts >= min_similarity                                                  # :91
            top, ts = top[keep], ts[keep]
            rows.extend([start + r] * len(top))
            cols.extend(top.tolist())
            sims.extend(ts.tolist())
    return np.asarray(rows, np.int32), np.asarray(cols, np.int32), np.asarray(sims, np.float32)


# --------------------------------------------------------------------------------------
# f2 / f3 (next rows): sampled evaluation and the optimizer step
# --------------------------------------------------------------------------------------


def eval_ranks(z: torch.Tensor, n_users: int, users: torch.Tensor, candidates: torch.Tensor):
    """Inner loop of eval_sampled, scripts/train_gat_custom.py:200-206: per evaluated user the scores of
    [positive] + negatives, ``i_emb @ u_emb``, and ``rank = (scores > scores[0]).sum() + 1``.  Returns (ranks, scores)."""
    U, I = z[:n_users], z[n_users:]
    scores = torch.einsum("qkc,qc->qk", I[candidates], U[users])
    ranks = (scores > scores[:, :1]).sum(dim=1) + 1
    return ranks, scores


def sample_eval_candidates(train_pos_idx, eval_pos, n_items: int, neg_k: int):
    """Candidate lists of eval_sampled, scripts/train_gat_custom.py:190-199: for each evaluated user, in dict order, the
    positive followed by ``neg_k`` negatives drawn one ``np.random.randint`` at a time and rejected while they are in the
    user's training positives or equal the positive.  Consumes the global numpy stream exactly like the reference."""
    user_pos_sets = {u: set(pos) for u, pos in train_pos_idx.items()}
    users, cands = [], []
    for u, pos_i in eval_pos.items():
        avoid = user_pos_sets.get(u, set()) | {pos_i}
        negs = []
        while len(negs) < neg_k:
            cand = np.random.randint(0, n_items)
            if cand not in avoid:
                negs.append(cand)
        users.append(u)
        cands.append([pos_i] + negs)
    return torch.tensor(users, dtype=torch.long), torch.tensor(cands, dtype=torch.long).reshape(len(users), neg_k + 1)


def ranking_metrics(ranks, Ks=(10, 20)):
    """scripts/train_gat_custom.py:206-210."""
    import math
    out = {f"recall@{k}": [] for k in Ks}
    out.update({f"ndcg@{k}": [] for k in Ks})
    for rank in ranks.tolist():
        for k in Ks:
            hit = 1.0 if rank <= k else 0.0
            out[f"recall@{k}"].append(hit)
            out[f"ndcg@{k}"].append((1.0 / math.log2(rank + 1)) if hit else 0.0)
    return {m: float(np.mean(v)) if v else 0.0 for m, v in out.items()}
