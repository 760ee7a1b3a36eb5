"""The callers either side of the hot path: the optimizer step and the sampled evaluator."""
from __future__ import annotations

from typing import Dict, Iterable, Sequence

import torch

from . import _lib


class Adam(torch.optim.Optimizer):
    """``torch.optim.Adam(params, lr, weight_decay)`` as the reference constructs it (scripts/train_gat_custom.py:335):
    same update rule (L2 in the gradient, bias correction, eps outside the square root), one fused kernel per tensor."""

    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("b200gat.Adam: parameters must be float32 CUDA tensors (there is no CPU fallback)")
                if not p.is_contiguous():
                    raise RuntimeError("b200gat.Adam: parameters must be contiguous")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] = int(st["step"]) + 1          # a torch.optim.Adam state_dict carries the step as a tensor
                g = p.grad.contiguous()
                with torch.cuda.device(p.device):
                    _lib.call("b200gat_adam_step_f32", _lib.ptr(p), _lib.ptr(g), _lib.ptr(st["exp_avg"]), _lib.ptr(st["exp_avg_sq"]),
                              p.numel(), group["lr"], b1, b2, group["eps"], group["weight_decay"], st["step"], _lib.stream())
        return loss


def sample_bpr_epoch(graph, n_users: int, n_items: int, samples: int, seed: int):
    """Device version of ``sample_bpr_epoch`` (scripts/train_gat_custom.py:213-224): ``samples`` triples (u, i, j) with u
    uniform over users that have training positives, i a uniform positive of u, j a uniform item u has not interacted with.
    ``graph`` is the :class:`GraphStructure` of the training ``edge_index`` (a user's positives are its out-edges).  Returns
    int64 CUDA tensors; same distribution as the reference, not the same stream (Python's ``random`` cannot be replayed)."""
    dev = graph.colptr.device
    u = torch.empty(samples, dtype=torch.int64, device=dev)
    i = torch.empty_like(u)
    j = torch.empty_like(u)
    n_fail = torch.empty(1, dtype=torch.int32, device=dev)
    # the users that have positives (the keys of the reference's user -> positives dict), once per graph
    active = getattr(graph, "_active_users", None)
    if active is None:
        deg = graph.colptr[1:n_users + 1] - graph.colptr[:n_users]
        active = torch.nonzero(deg > 0).flatten().to(torch.int32).contiguous()
        graph._active_users = active
    if active.numel() == 0:
        raise RuntimeError("sample_bpr_epoch: no user has a training positive")
    with torch.cuda.device(dev):
        _lib.call("b200gat_sample_bpr_ex", _lib.ptr(graph.colptr), _lib.ptr(graph.row), n_users, n_items, samples, int(seed) & (2 ** 64 - 1),
                  _lib.ptr(active), int(active.numel()), _lib.ptr(u), _lib.ptr(i), _lib.ptr(j), _lib.ptr(n_fail), _lib.stream())
    if int(n_fail.item()):
        raise RuntimeError("sample_bpr_epoch: some sampled user has interacted with every item (no negative exists)")
    return u, i, j


def eval_ranks(z: torch.Tensor, n_users: int, users: torch.Tensor, candidates: torch.Tensor) -> torch.Tensor:
    """ranks[q] = (scores > scores[0]).sum() + 1 for scores = I[candidates[q]] @ U[users[q]]
    (scripts/train_gat_custom.py:200-206).  ``candidates`` [n_eval, 1+K] int64 item ids, column 0 = the positive."""
    if not z.is_cuda:
        raise RuntimeError("b200gat eval_ranks: z must be a CUDA tensor (there is no CPU fallback)")
    z = _lib._f32(z.detach(), "z").contiguous()
    n, c = z.shape
    users = users.to(z.device, torch.int64).contiguous()
    candidates = candidates.to(z.device, torch.int64).contiguous()
    if candidates.dim() != 2 or candidates.shape[0] != users.shape[0]:
        raise RuntimeError("candidates must be [n_eval, 1+K]")
    n_eval, k1 = candidates.shape
    ranks = torch.empty(n_eval, dtype=torch.int32, device=z.device)
    n_bad = torch.empty(1, dtype=torch.int32, device=z.device)
    with torch.cuda.device(z.device):
        _lib.call("b200gat_eval_ranks_f32", _lib.ptr(z), n_users, n - n_users, c, _lib.ptr(users), _lib.ptr(candidates), n_eval, k1,
                  _lib.ptr(ranks), _lib.ptr(n_bad), _lib.stream())
    if int(n_bad.item()):
        raise IndexError("eval_ranks: user or candidate id out of range")
    return ranks


def ranking_metrics(ranks: torch.Tensor, Ks: Sequence[int] = (10, 20)) -> Dict[str, float]:
    """Recall@k / NDCG@k from ranks, as scripts/train_gat_custom.py:206-210 (mean over evaluated users, 0.0 if none)."""
    out = {}
    r = ranks.to(torch.float64)
    for k in Ks:
        hit = (r <= k).to(torch.float64)
        out[f"recall@{k}"] = float(hit.mean()) if r.numel() else 0.0
    for k in Ks:
        hit = (r <= k).to(torch.float64)
        out[f"ndcg@{k}"] = float((hit / torch.log2(r + 1)).mean()) if r.numel() else 0.0
    return out


def eval_sampled(model, item_feats: torch.Tensor, edge_index: torch.Tensor, users: torch.Tensor, candidates: torch.Tensor,
                 Ks: Sequence[int] = (10, 20)) -> Dict[str, float]:
    """One no_grad forward + ranks + metrics: the device part of eval_sampled (scripts/train_gat_custom.py:184-210).  The
    negatives are an input: the reference draws them with numpy on the host (:193-198), which stays host code."""
    with torch.no_grad():
        z = model(item_feats, edge_index)
    return ranking_metrics(eval_ranks(z, model.n_users, users, candidates), Ks)
