"""torch.autograd ops over the C ABI: one GAT layer (both dialects) and the ranking loss."""
from __future__ import annotations

import torch

from . import _lib
from .graph import GraphStructure


def _empty(shape, like, dtype=torch.float32):  # noqa: D401
    return torch.empty(shape, dtype=dtype, device=like.device)


class GATLayerFunction(torch.autograd.Function):
    """y = GAT(x; W, a_src, a_dst[, bias]) over ``graph``.

    policy CUSTOM = SimpleGATLayer.forward (scripts/train_gat_custom.py:75-93), policy PYG =
    GATConv(heads=H, concat=False, add_self_loops=False) (scripts/train_gat_pyg.py:77,87).
    Saved for backward: x, h, the per-node scalars (s, rowstat) and the output -- no E-sized float tensor.
    """

    @staticmethod
    def forward(ctx, x, weight, a_src, a_dst, bias, graph: GraphStructure, heads: int, channels: int, policy: int,
                negative_slope: float, p_drop: float, seed: int, bf16: bool = False):
        if not x.is_cuda:
            raise RuntimeError("b200gat GAT layer: x must be a CUDA tensor (there is no CPU fallback)")
        x = _lib._f32(x, "x").contiguous()
        weight = _lib._f32(weight, "weight").contiguous()
        a_s = _lib._f32(a_src, "a_src").contiguous().view(heads, channels)
        a_d = _lib._f32(a_dst, "a_dst").contiguous().view(heads, channels)
        n, f_in = x.shape
        if n != graph.n_nodes:
            raise RuntimeError(f"x has {n} rows but the graph has {graph.n_nodes} nodes")
        if weight.shape != (heads * channels, f_in):
            raise RuntimeError(f"weight must be [{heads * channels}, {f_in}], got {tuple(weight.shape)}")
        need_grad = any(ctx.needs_input_grad[:5])
        with torch.cuda.device(x.device):
            st = _lib.stream()
            h = _empty((n, heads * channels), x, torch.bfloat16 if bf16 else torch.float32)
            s = _empty((n, 2 * heads), x)
            ws_bytes = _lib.dense_workspace_bytes(heads, channels, f_in)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            _lib.call("b200gat_project_bf16" if bf16 else "b200gat_project_f32", _lib.ptr(x), _lib.ptr(weight), _lib.ptr(a_s),
                      _lib.ptr(a_d), n, f_in, heads, channels, _lib.ptr(h), _lib.ptr(s), _lib.ptr(ws), ws_bytes, st)
            out = _empty((n, channels), x)
            rowstat = _empty((n, heads, 2), x) if need_grad else None
            out_heads = _empty((n, heads, channels), x) if (need_grad and heads > 1) else None
            b = None if bias is None else _lib._f32(bias, "bias").contiguous()
            sf = graph.sched_fwd
            _lib.call("b200gat_edge_fwd_bf16" if bf16 else "b200gat_edge_fwd_f32", _lib.ptr(h), _lib.ptr(s), _lib.ptr(sf.sched),
                      sf.n_sched, _lib.ptr(sf.table),
                      sf.n_long, _lib.ptr(sf.partial(heads * (channels + 4))), _lib.ptr(graph.col), _lib.ptr(graph.perm), 0,
                      heads, channels, policy, negative_slope, _lib.ptr(b), _lib.ptr(out), _lib.ptr(out_heads),
                      _lib.ptr(rowstat), p_drop, seed, st)
        if need_grad:
            ctx.save_for_backward(x, weight, a_s, a_d, b, h, s, rowstat, out if heads == 1 else out_heads)
            ctx.graph = graph
            ctx.cfg = (heads, channels, policy, negative_slope, p_drop, seed, bf16)
            ctx.att_shape = (a_src.shape, a_dst.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, a_s, a_d, b, h, s, rowstat, out_h = ctx.saved_tensors
        heads, channels, policy, negative_slope, p_drop, seed, bf16 = ctx.cfg
        g = ctx.graph
        n, f_in = x.shape
        dout = dout.contiguous()
        with torch.cuda.device(x.device):
            st = _lib.stream()
            nodestat = _empty((n, heads, 4), x)
            ws_bytes = _lib.dense_workspace_bytes(heads, channels, f_in)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            dbias = torch.empty_like(b) if (b is not None and ctx.needs_input_grad[4]) else None
            dout_g = _empty((n, channels), x, torch.bfloat16) if bf16 else None   # bf16 copy that the edge kernel gathers
            _lib.call("b200gat_node_prep_f32", _lib.ptr(dout), _lib.ptr(out_h), _lib.ptr(b if heads == 1 else None),
                      _lib.ptr(s), _lib.ptr(rowstat), n, 0, heads, channels, _lib.ptr(nodestat), _lib.ptr(dbias),
                      _lib.ptr(dout_g), _lib.ptr(ws), ws_bytes, st)
            dh = _empty((n, heads * channels), x)
            de = _empty((max(g.n_edges, 1), heads), x)
            ds = _empty((n, 2 * heads), x)
            sb = g.sched_bwd
            _lib.call("b200gat_edge_bwd_bf16" if bf16 else "b200gat_edge_bwd_f32", _lib.ptr(h), _lib.ptr(s),
                      _lib.ptr(dout_g if bf16 else dout), _lib.ptr(nodestat),
                      _lib.ptr(sb.sched), sb.n_sched, _lib.ptr(sb.table), sb.n_long,
                      _lib.ptr(sb.partial(heads * channels + 4)), _lib.ptr(g.row), _lib.ptr(g.perm_csc), 0, heads, channels,
                      policy, negative_slope, _lib.ptr(dh), _lib.ptr(de), _lib.ptr(ds), 2 * heads, p_drop, seed, st)
            _lib.call("b200gat_ds_dst_f32", _lib.ptr(de), _lib.ptr(g.rowptr), _lib.ptr(g.csr2csc), n, g.n_edges, heads,
                      _lib.ptr(ds, heads), 2 * heads, st)
            del de
            dx = _empty((n, f_in), x) if ctx.needs_input_grad[0] else None
            dw = torch.empty_like(weight)
            da_s = torch.empty_like(a_s)
            da_d = torch.empty_like(a_d)
            _lib.call("b200gat_project_bwd_bf16" if bf16 else "b200gat_project_bwd_f32", _lib.ptr(x), _lib.ptr(weight), _lib.ptr(a_s), _lib.ptr(a_d), _lib.ptr(dh),
                      _lib.ptr(ds), n, f_in, heads, channels, _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(da_s), _lib.ptr(da_d),
                      _lib.ptr(ws), ws_bytes, st)
        sa, sd = ctx.att_shape
        return dx, dw, da_s.view(sa), da_d.view(sd), dbias, None, None, None, None, None, None, None, None


class NodeFeaturesFunction(torch.autograd.Function):
    """x0 = cat[user_emb.weight, item_proj(item_feats)] (CustomGAT.node_features, scripts/train_gat_custom.py:105-109)
    without the concat copy: the projection writes straight into the tail rows of the [N, C] buffer."""

    @staticmethod
    def forward(ctx, user_w, proj_w, proj_b, item_feats, tensor_core=False):
        if not item_feats.is_cuda:
            raise RuntimeError("b200gat node_features: tensors must be CUDA tensors (there is no CPU fallback)")
        nu, c = user_w.shape
        ni, f = item_feats.shape
        feats = _lib._f32(item_feats, "item_feats").contiguous()
        w = _lib._f32(proj_w, "item_proj.weight").contiguous()
        b = None if proj_b is None else _lib._f32(proj_b, "item_proj.bias").contiguous()
        x0 = _empty((nu + ni, c), feats)
        x0[:nu].copy_(user_w)
        ws_bytes = _lib.dense_workspace_bytes(1, c, f)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=feats.device)
        with torch.cuda.device(feats.device):
            _lib.call("b200gat_linear_tc_f32" if tensor_core else "b200gat_linear_f32", _lib.ptr(feats), _lib.ptr(w), _lib.ptr(b), ni, f, c,
                      _lib.ptr(x0, nu * c), c, _lib.ptr(ws), ws_bytes, _lib.stream())
        ctx.save_for_backward(feats)
        ctx.dims = (nu, ni, c, f, b is not None)
        return x0

    @staticmethod
    def backward(ctx, dx0):
        (feats,) = ctx.saved_tensors
        nu, ni, c, f, has_bias = ctx.dims
        dx0 = dx0.contiguous()
        dw = torch.empty((c, f), dtype=torch.float32, device=feats.device)
        db = torch.empty((c,), dtype=torch.float32, device=feats.device) if has_bias else None
        ws_bytes = _lib.dense_workspace_bytes(1, c, f)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=feats.device)
        with torch.cuda.device(feats.device):
            _lib.call("b200gat_linear_bwd_f32", _lib.ptr(feats), _lib.ptr(dx0, nu * c), c, ni, f, c, _lib.ptr(dw), _lib.ptr(db),
                      _lib.ptr(ws), ws_bytes, _lib.stream())
        return dx0[:nu], dw, db, None, None


def node_features(user_w, proj_w, proj_b, item_feats, tensor_core=False):
    """``tensor_core``: project the item features with the TF32-split tensor-core kernel (the bf16 tier) instead of fp32 FFMA."""
    if item_feats.requires_grad:
        raise NotImplementedError("b200gat node_features: item_feats is an input and must not require grad")
    return NodeFeaturesFunction.apply(user_w, proj_w, proj_b, item_feats, bool(tensor_core))


def gat_layer(x, weight, a_src, a_dst, bias, graph, heads, channels, policy, negative_slope=0.2, p_drop=0.0, seed=0,
              feature_dtype=torch.float32):
    """``feature_dtype=torch.bfloat16`` selects the bf16 projection: h (and the dout copy the backward gathers) are stored as
    bf16, everything is accumulated and returned in fp32 (tolerance tier 2e-2 instead of 1e-5)."""
    if feature_dtype not in (torch.float32, torch.bfloat16):
        raise NotImplementedError(f"feature_dtype {feature_dtype} is not supported (float32 or bfloat16)")
    return GATLayerFunction.apply(x, weight, a_src, a_dst, bias, graph, heads, channels, policy, float(negative_slope),
                                  float(p_drop), int(seed), feature_dtype == torch.bfloat16)


class RankLossFunction(torch.autograd.Function):
    """Fused pos/neg dot products + BPR / BCE mean (scripts/train_gat_custom.py:350-359)."""

    @staticmethod
    def forward(ctx, z, n_users: int, u, i, j, kind: int):
        if not z.is_cuda:
            raise RuntimeError("b200gat ranking loss: z must be a CUDA tensor (there is no CPU fallback)")
        z = _lib._f32(z, "z").contiguous()
        n, c = z.shape
        n_items = n - n_users
        idx = []
        for name, t in (("u", u), ("i", i), ("j", j)):
            if t.dtype != torch.int64 or t.dim() != 1 or t.shape != u.shape:
                raise RuntimeError(f"{name} must be an int64 vector of the common length")
            idx.append(t.to(z.device).contiguous())
        s = int(u.shape[0])
        need = ctx.needs_input_grad[0]
        ws_bytes = _lib.loss_workspace_bytes(n, s)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=z.device)
        loss = _empty((1,), z)
        with torch.cuda.device(z.device):
            _lib.call("b200gat_rank_loss_fwd_f32", _lib.ptr(z), n_users, n_items, c, _lib.ptr(idx[0]), _lib.ptr(idx[1]),
                      _lib.ptr(idx[2]), s, None, kind, int(need), _lib.ptr(loss), _lib.ptr(ws), ws_bytes, _lib.stream())
        if need:
            ctx.save_for_backward(z, idx[0], idx[1], idx[2], ws)
            ctx.meta = (n_users, n_items, c, s, kind, ws_bytes)
        return loss.view(())

    @staticmethod
    def backward(ctx, grad_out):
        z, u, i, j, ws = ctx.saved_tensors
        n_users, n_items, c, s, kind, ws_bytes = ctx.meta
        go = grad_out.to(torch.float32).reshape(1).contiguous()
        dz = torch.empty_like(z)
        with torch.cuda.device(z.device):
            _lib.call("b200gat_rank_loss_bwd_f32", _lib.ptr(z), n_users, n_items, c, _lib.ptr(u), _lib.ptr(i), _lib.ptr(j), s,
                      None, kind, _lib.ptr(go), None, 0, n_users + n_items, _lib.ptr(dz), None, _lib.ptr(ws), ws_bytes,
                      _lib.stream())
        return dz, None, None, None, None, None


def bpr_loss(z: torch.Tensor, n_users: int, u: torch.Tensor, i: torch.Tensor, j: torch.Tensor) -> torch.Tensor:
    """``-log(sigmoid(pos - neg) + 1e-8).mean()`` with pos/neg as at scripts/train_gat_custom.py:350-355."""
    return RankLossFunction.apply(z, int(n_users), u, i, j, _lib.LOSS_BPR)


def bce_loss(z: torch.Tensor, n_users: int, u: torch.Tensor, i: torch.Tensor, j: torch.Tensor) -> torch.Tensor:
    """BCE-with-logits over cat[pos, neg] vs cat[1, 0] (scripts/train_gat_custom.py:356-359)."""
    return RankLossFunction.apply(z, int(n_users), u, i, j, _lib.LOSS_BCE)
