#!/usr/bin/env python3
"""The reference's training loop (scripts/train_gat_custom.py:337-381) on a synthetic graph of its shape, with every device
step taken from b200gat: union graph (U-I interactions + GPU cosine kNN), drop-in model, device BPR sampler, fused loss,
Adam kernel, sampled evaluation, best-checkpoint save in the reference's ``{"state_dict", "config"}`` format.

    python examples/train_synthetic.py --kind custom --epochs 5            # needs a B200; there is no CPU path

Not part of the test suite; the pieces it composes are (tests/test_gpu_parity.py, tests/test_gpu_knn.py).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gat  # noqa: E402


def synthetic_interactions(n_users, n_items, n_inter, rng):
    """train_pos_idx / val_pos as the reference's load code produces them: per user an array of item ids, one held out."""
    users = rng.integers(0, n_users, size=n_inter)
    pop = 1.0 / (np.arange(1, n_items + 1) + 40.0) ** 0.9
    items = rng.permutation(n_items)[rng.choice(n_items, size=n_inter, p=pop / pop.sum())]
    train_pos, val_pos = {}, {}
    order = np.argsort(users, kind="stable")
    bounds = np.flatnonzero(np.diff(users[order])) + 1
    for chunk in np.split(order, bounds):
        u = int(users[chunk[0]])
        its = items[chunk]
        if len(its) >= 2:
            val_pos[u] = int(its[-1])
            its = its[:-1]
        train_pos[u] = its
    return train_pos, val_pos


def eval_candidates(train_pos, eval_pos, n_items, neg_k):
    """Negative sampling of eval_sampled (scripts/train_gat_custom.py:190-199): host numpy, as in the reference."""
    users, cands = [], []
    for u, pos in eval_pos.items():
        seen = set(train_pos.get(u, ()).tolist()) | {pos}
        neg = []
        while len(neg) < neg_k:
            j = int(np.random.randint(0, n_items))
            if j not in seen:
                neg.append(j)
        users.append(u)
        cands.append([pos] + neg)
    return torch.tensor(users, dtype=torch.int64), torch.tensor(cands, dtype=torch.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="custom", choices=["custom", "pyg"])
    ap.add_argument("--users", type=int, default=10_000)
    ap.add_argument("--items", type=int, default=20_000)
    ap.add_argument("--interactions", type=int, default=200_000)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--heads", type=int, default=1)
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--samples", type=int, default=200_000)
    ap.add_argument("--neg-k", type=int, default=100)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--out", default="gat_synthetic.pt")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        sys.exit("b200gat needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    nu, ni = args.users, args.items

    train_pos, val_pos = synthetic_interactions(nu, ni, args.interactions, rng)
    feats = torch.nn.functional.normalize(torch.randn(ni, 128), dim=1).to(dev)           # fused 128-d item features
    rows, cols, _ = b200gat.build_ii_knn(feats, k=20, min_similarity=-1.0)                # graphs/build_ii_knn.py on the GPU
    edge_index = b200gat.union_edge_index(b200gat.build_edge_index(nu, ni, train_pos), nu, rows.cpu(), cols.cpu()).to(dev)
    graph = b200gat.graph_for(edge_index, nu + ni)                                        # CSR/CSC, built once
    print(f"graph: {nu + ni} nodes, {edge_index.shape[1]} edges")

    if args.kind == "custom":
        model = b200gat.CustomGAT(nu, ni, 128, args.hidden, args.layers).to(dev)
    else:
        model = b200gat.PyGGAT(nu, ni, 128, args.hidden, args.layers, args.heads, 0.1).to(dev)
    opt = b200gat.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    users, cands = eval_candidates(train_pos, val_pos, ni, args.neg_k)
    best = -1.0
    for epoch in range(1, args.epochs + 1):
        model.train()
        u, i, j = b200gat.sample_bpr_epoch(graph, nu, ni, args.samples, seed=args.seed + epoch)
        z = model(feats, edge_index)
        loss = b200gat.bpr_loss(z, nu, u, i, j)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        model.eval()
        m = b200gat.eval_sampled(model, feats, edge_index, users, cands)
        print(f"epoch {epoch}: loss {loss.item():.4f}  val recall@20 {m['recall@20']:.4f}  ndcg@20 {m['ndcg@20']:.4f}")
        if m["ndcg@20"] > best:
            best = m["ndcg@20"]
            torch.save({"state_dict": model.state_dict(), "config": vars(args)}, args.out)


if __name__ == "__main__":
    main()
