"""Drop-in torch modules for the reference's two layer APIs and the two models that call them.

* :class:`SimpleGATLayer` / :class:`CustomGAT` mirror scripts/train_gat_custom.py:63-115
* :class:`GATConv` / :class:`PyGGAT` mirror the PyG call made at scripts/train_gat_pyg.py:68-88

Constructor signatures, parameter names, shapes and initialisers are the reference's, so a ``state_dict`` saved by the
reference trainer (``torch.save({"state_dict", "config"})``, train_gat_custom.py:376) loads here and vice versa (PyG
dialect: checkpoints are written with the PyG >= 2.5 key ``lin.weight``; the PyG <= 2.4 keys ``lin_src`` / ``lin_dst`` are
accepted on load only).
Only the layer forward/backward changes: it runs the sm_100a kernels behind the C ABI.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from .functional import gat_layer, node_features
from .graph import graph_for


def _dropout_seed(training: bool, p: float) -> int:
    if not training or p <= 0.0:
        return 0
    # drawn from torch's CPU generator: follows torch.manual_seed, costs no device sync
    return int(torch.randint(0, 2 ** 62, (1,)).item())


class SimpleGATLayer(torch.nn.Module):
    """Single-head GAT layer with the reference's custom softmax dialect (train_gat_custom.py:63-93)."""

    def __init__(self, in_dim: int, out_dim: int, attn_dropout: float = 0.1, feature_dtype=torch.float32):
        super().__init__()
        self.feature_dtype = feature_dtype      # torch.bfloat16: bf16 projection / bf16 gathers (not in the reference API)
        self.lin = torch.nn.Linear(in_dim, out_dim, bias=False)
        self.a_src = torch.nn.Parameter(torch.empty(out_dim))
        self.a_dst = torch.nn.Parameter(torch.empty(out_dim))
        torch.nn.init.xavier_uniform_(self.lin.weight)
        torch.nn.init.xavier_uniform_(self.a_src.unsqueeze(0))
        torch.nn.init.xavier_uniform_(self.a_dst.unsqueeze(0))
        self.out_dim = out_dim
        self.attn_dropout = float(attn_dropout)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        graph = graph_for(edge_index, x.shape[0])
        p = self.attn_dropout if self.training else 0.0
        return gat_layer(x, self.lin.weight, self.a_src, self.a_dst, None, graph, 1, self.out_dim, _lib.POLICY_CUSTOM,
                         0.2, p, _dropout_seed(self.training, p), self.feature_dtype)


class GATConv(torch.nn.Module):
    """The subset of ``torch_geometric.nn.GATConv`` that the reference uses:
    ``GATConv(hidden, hidden, heads=H, dropout=p, add_self_loops=False, concat=False)(x, edge_index)``.
    Anything outside that subset raises ``NotImplementedError`` instead of silently differing."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True, edge_dim=None,
                 fill_value="mean", bias: bool = True, residual: bool = False, feature_dtype=torch.float32, **kwargs):
        super().__init__()
        self.feature_dtype = feature_dtype      # torch.bfloat16: bf16 projection / bf16 gathers (not a PyG argument)
        if kwargs:
            raise NotImplementedError(f"GATConv: unsupported arguments {sorted(kwargs)}")
        if not isinstance(in_channels, int):
            raise NotImplementedError("GATConv: bipartite (tuple) in_channels is not supported")
        if concat:
            raise NotImplementedError("GATConv: concat=True is not supported (the reference uses concat=False)")
        if add_self_loops:
            raise NotImplementedError("GATConv: add_self_loops=True is not supported (the reference passes False)")
        if edge_dim is not None:
            raise NotImplementedError("GATConv: edge_dim / edge_attr is not supported")
        if residual:
            raise NotImplementedError("GATConv: residual=True is not supported")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, float(negative_slope), float(dropout)
        self.add_self_loops, self.edge_dim, self.fill_value, self.residual = add_self_loops, edge_dim, fill_value, residual
        self.lin = torch.nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self._renamed = None
        self.register_load_state_dict_post_hook(GATConv._restore_renamed)
        self.reset_parameters()

    @staticmethod
    def _glorot(t: torch.Tensor) -> None:
        bound = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
        with torch.no_grad():
            t.uniform_(-bound, bound)

    def reset_parameters(self) -> None:
        self._glorot(self.lin.weight)
        self._glorot(self.att_src)
        self._glorot(self.att_dst)
        if self.bias is not None:
            torch.nn.init.zeros_(self.bias)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # PyG <= 2.4 stored the shared projection as lin_src / lin_dst (loading only: a checkpoint saved from here carries
        # lin.weight, the PyG >= 2.5 name).  The child Linear reads the SAME dict object after this returns, so the alias is
        # added in place and taken out again by the post hook: the caller's checkpoint dict is unchanged after the load.
        old, dst, new = prefix + "lin_src.weight", prefix + "lin_dst.weight", prefix + "lin.weight"
        if old in state_dict and new not in state_dict:
            self._renamed = (state_dict, new, [(k, state_dict.pop(k)) for k in (old, dst) if k in state_dict])
            state_dict[new] = self._renamed[2][0][1]
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    @staticmethod
    def _restore_renamed(module, _incompatible) -> None:
        pending, module._renamed = getattr(module, "_renamed", None), None
        if pending is not None:
            sd, new, olds = pending
            sd.pop(new, None)
            for k, v in olds:
                sd[k] = v

    def forward(self, x, edge_index, edge_attr=None, size=None, return_attention_weights=None):
        if edge_attr is not None or size is not None or return_attention_weights is not None:
            raise NotImplementedError("GATConv.forward: edge_attr / size / return_attention_weights are not supported")
        if not isinstance(x, torch.Tensor) or not isinstance(edge_index, torch.Tensor):
            raise NotImplementedError("GATConv.forward: only Tensor x and Tensor edge_index are supported")
        graph = graph_for(edge_index, x.shape[0])
        p = self.dropout if self.training else 0.0
        return gat_layer(x, self.lin.weight, self.att_src, self.att_dst, self.bias, graph, self.heads, self.out_channels,
                         _lib.POLICY_PYG, self.negative_slope, p, _dropout_seed(self.training, p), self.feature_dtype)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})"


class CustomGAT(torch.nn.Module):
    """scripts/train_gat_custom.py:96-115 with the layers swapped for the B200 ones."""

    def __init__(self, n_users: int, n_items: int, item_feat_dim: int, hidden: int, layers: int, feature_dtype=torch.float32):
        super().__init__()
        self.n_users, self.n_items = n_users, n_items
        self.feature_dtype = feature_dtype
        self.user_emb = torch.nn.Embedding(n_users, hidden)
        torch.nn.init.normal_(self.user_emb.weight, std=0.1)
        self.item_proj = torch.nn.Linear(item_feat_dim, hidden)
        self.layers = torch.nn.ModuleList([SimpleGATLayer(hidden, hidden, feature_dtype=feature_dtype) for _ in range(layers)])

    def node_features(self, item_feats: torch.Tensor) -> torch.Tensor:
        # == torch.cat([self.user_emb.weight, self.item_proj(item_feats)], dim=0), without the concat copy
        return node_features(self.user_emb.weight, self.item_proj.weight, self.item_proj.bias, item_feats,
                             tensor_core=self.feature_dtype == torch.bfloat16)

    def forward(self, item_feats: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        x = self.node_features(item_feats)
        for gat in self.layers:
            x = gat(x, edge_index)
        return x


class PyGGAT(torch.nn.Module):
    """scripts/train_gat_pyg.py:68-88 with ``GATConv`` swapped for the B200 one."""

    def __init__(self, n_users: int, n_items: int, item_feat_dim: int, hidden: int, layers: int, heads: int,
                 attn_dropout: float, feature_dtype=torch.float32):
        super().__init__()
        self.n_users, self.n_items = n_users, n_items
        self.feature_dtype = feature_dtype
        self.user_emb = torch.nn.Embedding(n_users, hidden)
        torch.nn.init.normal_(self.user_emb.weight, std=0.1)
        self.item_proj = torch.nn.Linear(item_feat_dim, hidden)
        self.convs = torch.nn.ModuleList()
        for _ in range(layers):
            self.convs.append(GATConv(hidden, hidden, heads=heads, dropout=attn_dropout, add_self_loops=False,
                                      concat=False, feature_dtype=feature_dtype))

    def node_features(self, item_feats: torch.Tensor) -> torch.Tensor:
        # == torch.cat([self.user_emb.weight, self.item_proj(item_feats)], dim=0), without the concat copy
        return node_features(self.user_emb.weight, self.item_proj.weight, self.item_proj.bias, item_feats,
                             tensor_core=self.feature_dtype == torch.bfloat16)

    def forward(self, item_feats: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        x = self.node_features(item_feats)
        for conv in self.convs:
            x = conv(x, edge_index)
        return x
