"""b200gat -- the B200 (sm_100a) GAT training hot path of PlotPointe-GAT-Recommendation.

Drop-in modules (:class:`SimpleGATLayer`, :class:`GATConv`, :class:`CustomGAT`, :class:`PyGGAT`), the fused
ranking losses (:func:`bpr_loss`, :func:`bce_loss`) and the device graph builder, all backed by hand-written CUDA
behind the C ABI in ``include/b200gat.h``.  Importing this package fails loudly if ``libb200gat.so`` has not been
built; there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from .functional import bce_loss, bpr_loss, gat_layer
from .graph import GraphStructure, build_edge_index, build_graph, clear_graph_cache, graph_for, union_edge_index
from .knn import build_ii_knn, knn_neighbors
from .modules import CustomGAT, GATConv, PyGGAT, SimpleGATLayer
from .train import Adam, eval_ranks, eval_sampled, ranking_metrics, sample_bpr_epoch

__all__ = ["SimpleGATLayer", "GATConv", "CustomGAT", "PyGGAT", "bpr_loss", "bce_loss", "gat_layer", "build_graph",
           "graph_for", "build_edge_index", "GraphStructure", "clear_graph_cache", "Adam", "eval_ranks", "eval_sampled",
           "ranking_metrics", "build_ii_knn", "knn_neighbors", "sample_bpr_epoch", "union_edge_index"]
