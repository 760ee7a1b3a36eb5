"""ctypes binding of libb200gat.so (the C ABI declared in include/b200gat.h).

PyTorch is used for device memory and streams only: every call below hands raw device pointers, sizes
and the current CUDA stream to the library.  There is no CPU fallback -- if the shared library is
missing this module raises at import, and every wrapper rejects non-CUDA tensors.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200GAT_LIB", os.path.join(_HERE, "libb200gat.so"))   # override: A/B builds of the library

POLICY_CUSTOM, POLICY_PYG = 0, 1
LOSS_BPR, LOSS_BCE = 0, 1

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: the CUDA library has not been built. Run "
        "`python -c 'import __graft_entry__ as g; g.build()'` (or plotpointe-gat-recommendation_b200/csrc/build.sh). "
        "There is no CPU fallback for this package.")

_lib = ctypes.CDLL(LIB_PATH)

_P = c_void_p
_SIGS = {
    "b200gat_last_error": (ctypes.c_char_p, []),
    "b200gat_abi_version": (c_int, []),
    "b200gat_launch_count": (c_int64, []),
    "b200gat_graph_workspace_bytes": (c_int, [c_int64, c_int64, ctypes.POINTER(c_size_t)]),
    "b200gat_build_graph": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "b200gat_set_gemm_mode": (c_int, [c_int]),
    "b200gat_get_gemm_mode": (c_int, []),
    "b200gat_project_f32": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "b200gat_dense_workspace_bytes": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "b200gat_project_bwd_f32": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, _P,
                                        c_size_t, _P]),
    "b200gat_linear_f32": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, c_int64, _P, c_size_t, _P]),
    "b200gat_linear_tc_f32": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, c_int64, _P, c_size_t, _P]),
    "b200gat_linear_bwd_f32": (c_int, [_P, _P, c_int64, c_int64, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "b200gat_adam_step_f32": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int64, _P]),
    "b200gat_sample_bpr": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_uint64, _P, _P, _P, _P, _P]),
    "b200gat_sample_bpr_ex": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_uint64, _P, c_int64, _P, _P, _P, _P, _P]),
    "b200gat_eval_ranks_f32": (c_int, [_P, c_int64, c_int64, c_int, _P, _P, c_int64, c_int, _P, _P, _P]),
    "b200gat_knn_workspace_bytes": (c_int, [c_int64, c_int, ctypes.POINTER(c_size_t)]),
    "b200gat_knn_cosine_f32": (c_int, [_P, c_int64, c_int, c_int, c_float, _P, _P, _P, _P, _P, c_size_t, _P]),
    "b200gat_colsum_f32": (c_int, [_P, c_int64, c_int, _P, _P, c_size_t, _P]),
    "b200gat_edge_fwd_f32": (c_int, [_P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int, c_int, c_int, c_float, _P,
                                     _P, _P, _P, c_float, c_uint64, _P]),
    "b200gat_node_prep_f32": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "b200gat_project_bf16": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "b200gat_project_bwd_bf16": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, _P,
                                         c_size_t, _P]),
    "b200gat_edge_fwd_bf16": (c_int, [_P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int, c_int, c_int, c_float, _P,
                                      _P, _P, _P, c_float, c_uint64, _P]),
    "b200gat_edge_bwd_bf16": (c_int, [_P, _P, _P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int, c_int, c_int,
                                      c_float, _P, _P, _P, c_int, c_float, c_uint64, _P]),
    "b200gat_project_bf16_ex": (c_int, [_P, c_int, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "b200gat_project_bwd_bf16_ex": (c_int, [_P, c_int, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, c_int, _P, _P, _P, _P,
                                            c_size_t, _P]),
    "b200gat_edge_fwd_stream_f32": (c_int, [_P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int, c_int, c_float, _P, _P, _P,
                                            c_float, c_uint64, c_float, c_int, _P]),
    "b200gat_edge_fwd_stream_bf16": (c_int, [_P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int, c_int, c_float, _P, _P, _P,
                                             c_float, c_uint64, c_float, c_int, _P]),
    "b200gat_node_stat_f32": (c_int, [_P, _P, c_int64, c_int64, c_int, _P, _P]),
    "b200gat_node_stat_set_t_f32": (c_int, [_P, _P, c_int64, _P]),
    "b200gat_edge_bwd_phase1_f32": (c_int, [_P, _P, _P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int, c_int, c_float,
                                            _P, _P, _P, c_int, c_float, c_uint64, _P]),
    "b200gat_edge_bwd_phase1_bf16": (c_int, [_P, _P, _P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int, c_int, c_float,
                                             _P, _P, _P, c_int, c_float, c_uint64, _P]),
    "b200gat_edge_bwd_phase2_f32": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_float, _P, c_int, _P]),
    "b200gat_cast_bf16": (c_int, [_P, _P, c_int64, c_float, _P]),
    "b200gat_schedule_workspace_bytes": (c_int, [c_int64, ctypes.POINTER(c_size_t)]),
    "b200gat_build_schedule": (c_int, [_P, c_int64, c_int64, _P, _P, c_size_t, _P]),
    "b200gat_edge_bwd_f32": (c_int, [_P, _P, _P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int, c_int, c_int,
                                     c_float, _P, _P, _P, c_int, c_float, c_uint64, _P]),
    "b200gat_ds_dst_f32": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, _P, c_int, _P]),
    "b200gat_loss_workspace_bytes": (c_int, [c_int64, c_int64, ctypes.POINTER(c_size_t)]),
    "b200gat_rank_loss_fwd_f32": (c_int, [_P, c_int64, c_int64, c_int, _P, _P, _P, c_int64, _P, c_int, c_int, _P, _P,
                                          c_size_t, _P]),
    "b200gat_rank_loss_bwd_f32": (c_int, [_P, c_int64, c_int64, c_int, _P, _P, _P, c_int64, _P, c_int, _P, _P, c_int64,
                                          c_int64, _P, _P, _P, c_size_t, _P]),
    "b200gat_peer_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p)]),
    "b200gat_peer_free": (c_int, [_P]),
    "b200gat_peer_export": (c_int, [_P, _P, c_size_t]),
    "b200gat_peer_open": (c_int, [_P, ctypes.POINTER(c_void_p)]),
    "b200gat_peer_close": (c_int, [_P]),
    "b200gat_peer_pull": (c_int, [_P, _P, c_size_t, _P]),
    "b200gat_peer_flag_bytes": (c_int, [c_int, ctypes.POINTER(c_size_t)]),
    "b200gat_peer_signal": (c_int, [_P, c_int, c_int, c_int, ctypes.c_uint32, _P]),
    "b200gat_peer_wait": (c_int, [_P, c_int, c_int, c_int, ctypes.c_uint32, _P]),
    "b200gat_peer_allgather": (c_int, [_P, c_int, c_int, c_int, ctypes.c_uint32, c_int, _P, _P, _P]),
    "b200gat_project_push_f32": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, c_int, _P, c_size_t, _P]),
    "b200gat_project_bwd_push_f32": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, c_int, _P, _P, _P, _P,
                                             c_size_t, _P]),
    "b200gat_peer_push": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P]),
    "b200gat_peer_reduce_f32": (c_int, [_P, c_int, c_int, c_int, ctypes.c_uint32, c_uint64, c_int64, c_int64, _P, _P]),
    "b200gat_rank_loss_fwd_peer_f32": (c_int, [_P, c_int, c_int64, c_int64, c_int64, c_int, _P, _P, _P, c_int64, c_int64, c_int64,
                                               _P, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "b200gat_rank_loss_bwd_peer_f32": (c_int, [_P, c_int, c_int64, c_int64, c_int64, c_int, _P, _P, _P, c_int64, _P, c_int, _P,
                                               _P, _P, c_int64, _P, _P, _P, c_size_t, _P]),
}
EXPORTS = tuple(_SIGS)
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(_lib, _name)      # AttributeError here = header/library mismatch: fail at import
    _fn.restype = _res
    _fn.argtypes = _args

# kernels launched by each entry point (used for bench.py's gpu_launches claim; graph/loss sort passes are
# counted by the library itself, see b200gat_launch_count)
timing = None      # bench.py sets this to a dict: entry point -> list of (start_event, end_event)


def last_error() -> str:
    return _lib.b200gat_last_error().decode()


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def ptr(t, offset_elems: int = 0):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("b200gat: expected a CUDA tensor (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("b200gat: expected a contiguous tensor")
    return c_void_p(t.data_ptr() + offset_elems * t.element_size())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, name):
    if t.dtype != torch.float32:
        raise RuntimeError(f"b200gat: {name} must be float32, got {t.dtype}")
    return t


def call(name: str, *args) -> None:
    if timing is None:
        _check(getattr(_lib, name)(*args), name)
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _check(getattr(_lib, name)(*args), name)
    b.record()
    timing.setdefault(name, []).append((a, b))


GEMM_FP32, GEMM_TF32X3 = 0, 1


def set_gemm_mode(mode: int) -> None:
    """GEMM_TF32X3 (default): tcgen05 tensor cores with the 3-term TF32 split; GEMM_FP32: CUDA-core FFMA."""
    _check(_lib.b200gat_set_gemm_mode(mode), "set_gemm_mode")


def get_gemm_mode() -> int:
    return int(_lib.b200gat_get_gemm_mode())


def launch_count() -> int:
    """Total number of kernels this process has launched through the library."""
    return int(_lib.b200gat_launch_count())


def graph_workspace_bytes(n_nodes: int, n_edges: int) -> int:
    out = c_size_t(0)
    _check(_lib.b200gat_graph_workspace_bytes(n_nodes, n_edges, ctypes.byref(out)), "graph_workspace_bytes")
    return out.value


def schedule_workspace_bytes(n_rows: int) -> int:
    out = c_size_t(0)
    _check(_lib.b200gat_schedule_workspace_bytes(n_rows, ctypes.byref(out)), "schedule_workspace_bytes")
    return out.value


SPLIT_DEGREE = 128   # rows longer than this are cut into segments of this many edges (one warp each)


class Schedule:
    """Row schedule of one direction: ``sched`` [n_sched, 4] = (row, beg, end, slot+1 | 0) in descending-degree order
    with long rows cut into segments; ``table`` [n_long, 4] = (row, first slot, n segments, degree)."""

    def __init__(self, sched, table, n_slots):
        self.sched, self.table, self.n_slots = sched, table, n_slots
        self.n_sched = int(sched.shape[0]) if sched is not None else 0
        self.n_long = 0 if table is None else int(table.shape[0])

    def partial(self, floats_per_slot: int):
        if self.n_slots == 0:
            return None
        return torch.empty(self.n_slots * floats_per_slot, dtype=torch.float32, device=self.sched.device)


def make_schedule(ptr_tensor, offset: int, n_rows: int, degree_bound: int, split: int = None) -> Schedule:
    """Descending-degree schedule with rows of more than ``split`` edges cut into ``split``-edge segments (host logic,
    once per graph: a handful of torch ops on the sorted schedule)."""
    split = SPLIT_DEGREE if split is None else split
    sched = build_schedule(ptr_tensor, offset, n_rows, degree_bound)
    if n_rows == 0:
        return Schedule(sched[:0], None, 0)
    if os.environ.get("B200GAT_SCHED", "degree") == "natural" or split <= 0:
        return Schedule(sched, None, 0)
    deg = (sched[:, 2] - sched[:, 1]).long()
    n_long = int((deg > split).sum())            # the schedule is sorted: the long rows are the first n_long entries
    if n_long == 0:
        return Schedule(sched, None, 0)
    lng = sched[:n_long].long()
    nseg = (deg[:n_long] + split - 1) // split
    first = torch.cumsum(nseg, 0) - nseg
    total = int(nseg.sum())
    rid = torch.repeat_interleave(torch.arange(n_long, device=sched.device), nseg)
    k = torch.arange(total, device=sched.device) - first[rid]
    beg = lng[rid, 1] + k * split
    end = torch.minimum(lng[rid, 2], beg + split)
    segs = torch.stack([lng[rid, 0], beg, end, torch.arange(1, total + 1, device=sched.device)], dim=1).to(torch.int32)
    table = torch.stack([lng[:, 0], first, nseg, deg[:n_long]], dim=1).to(torch.int32).contiguous()
    return Schedule(torch.cat([segs, sched[n_long:]], dim=0).contiguous(), table, total)


def build_schedule(ptr_tensor, offset: int, n_rows: int, degree_bound: int):
    """Descending-degree row schedule (int32 [n_rows, 4]) of ptr_tensor[offset : offset+n_rows+1]."""
    if os.environ.get("B200GAT_SCHED", "degree") == "natural":
        degree_bound = 0
    sched = torch.empty((max(n_rows, 1), 4), dtype=torch.int32, device=ptr_tensor.device)
    wsb = schedule_workspace_bytes(n_rows)
    ws = torch.empty(wsb, dtype=torch.uint8, device=ptr_tensor.device)
    with torch.cuda.device(ptr_tensor.device):
        call("b200gat_build_schedule", ptr(ptr_tensor, offset), n_rows, degree_bound, ptr(sched), ptr(ws), wsb, stream())
    return sched


def dense_workspace_bytes(heads: int, channels: int, in_features: int) -> int:
    out = c_size_t(0)
    _check(_lib.b200gat_dense_workspace_bytes(heads, channels, in_features, ctypes.byref(out)), "dense_workspace_bytes")
    return out.value


def loss_workspace_bytes(n_nodes: int, n_triples: int) -> int:
    out = c_size_t(0)
    _check(_lib.b200gat_loss_workspace_bytes(n_nodes, n_triples, ctypes.byref(out)), "loss_workspace_bytes")
    return out.value
