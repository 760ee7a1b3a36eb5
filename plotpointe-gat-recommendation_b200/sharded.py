"""Row-sharded full-graph GAT training over P GPUs of one node (one process per GPU).

Nodes are dealt round-robin to the P ranks (user u -> rank u % P, item i -> rank i % P), which balances rows and, for
the reference's graphs, edges.  Rank p owns its nodes' rows: their input features, their destination rows in the
forward (CSR slice) and their source rows in the backward (CSC slice).

All exchanges of a step run on the device over peer-mapped memory (``PeerFabric``, b200gat_peer_* in include/b200gat.h):
every rank keeps one exported buffer with the same layout; a producer kernel writes this rank's row block in place, a
signal kernel raises a flag in every peer's buffer, and the consuming kernel waits on its own flags and then pulls the
peers' blocks over NVLink with all SMs.  No host round trip and no library collective inside a step; ``torch.distributed``
only carries the set-up (IPC handles) and the bench's timing reductions.

  forward, heads == 1 : all-gather of the projected rows [h | s]        (N x (C + 2) values)   -> fused edge forward
  forward, heads  > 1 : all-gather of the layer INPUT rows x (F_in wide, H times narrower than h); every rank projects all
                        rows itself (redundant GEMM work instead of H x the NVLink bytes)
  backward            : all-gather of dout rows and per-destination scalars (N x (C + 4H)) -> fused edge backward on the local
                        source rows (no atomics, no reduce-scatter); reduce-pull of the ds_dst partial sums (N x H) and of the
                        parameter gradients (< 100 KB), both summed in rank order on every rank (bitwise identical everywhere)
  loss                : the S triples are split over the ranks; a triple's three Z rows are read straight out of the owners'
                        blocks (S x 3 rows over NVLink instead of an all-gather of N rows); the per-triple coefficients
                        (2 S floats) are exchanged and each rank forms the gradient rows of its own nodes the same way.

The plan (row layout, per-rank edge selections) is plain torch and also runs on CPU tensors, which is how the gloo tests
exercise it.
"""
from __future__ import annotations

import ctypes
import json
import os
import statistics
import time
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------------ the plan
@dataclass
class ShardPlan:
    """Round-robin block layout.  Users u and items i go to rank u % P / i % P, so every rank owns (within one row) the
    same number of users and of items and, for a random graph, the same number of edges.  Inside a block users come
    first, then items (the order node_features produces).  Blocks are padded to ``n_max`` rows so that every rank's block has
    the same size; ``perm_map[v]`` is the row of node v in the gathered [P*n_max, ...] tensors, and every index array the
    kernels see (col, row, schedules) is expressed in that row space."""
    rank: int
    world: int
    n_users: int
    n_items: int
    n_max: int
    n_loc: int
    cu: int                      # users owned by this rank
    ci: int                      # items owned by this rank
    perm_map: torch.Tensor       # int64 [N]  node id -> gathered row
    local_nodes: torch.Tensor    # int64 [n_loc] node ids of the local rows, in local row order
    fwd_sel: torch.Tensor        # int64 ids (ascending) of the edges whose destination is local
    bwd_sel: torch.Tensor        # int64 ids (ascending) of the edges whose source is local

    @property
    def lo(self):
        return self.rank * self.n_max

    @property
    def n_rows_total(self):
        return self.world * self.n_max


def owner_of(nodes: torch.Tensor, n_users: int, world: int) -> torch.Tensor:
    return torch.where(nodes < n_users, nodes % world, (nodes - n_users) % world)


def make_plan(edge_index: torch.Tensor, n_users: int, n_items: int, rank: int, world: int) -> ShardPlan:
    dev = edge_index.device
    cu_all = [(n_users - b + world - 1) // world for b in range(world)]
    ci_all = [(n_items - b + world - 1) // world for b in range(world)]
    n_max = (max(a + b for a, b in zip(cu_all, ci_all)) + 3) // 4 * 4     # x4: every per-row block is a multiple of 16 bytes
    cu_t = torch.tensor(cu_all, dtype=torch.int64, device=dev)
    u = torch.arange(n_users, dtype=torch.int64, device=dev)
    i = torch.arange(n_items, dtype=torch.int64, device=dev)
    perm_map = torch.cat([(u % world) * n_max + u // world, (i % world) * n_max + cu_t[i % world] + i // world])
    local_nodes = torch.cat([torch.arange(rank, max(n_users, rank), world, dtype=torch.int64, device=dev),
                             n_users + torch.arange(rank, max(n_items, rank), world, dtype=torch.int64, device=dev)])
    src, dst = edge_index[0], edge_index[1]
    fwd_sel = torch.nonzero(owner_of(dst, n_users, world) == rank).flatten()
    bwd_sel = torch.nonzero(owner_of(src, n_users, world) == rank).flatten()
    return ShardPlan(rank, world, n_users, n_items, n_max, cu_all[rank] + ci_all[rank], cu_all[rank], ci_all[rank], perm_map,
                     local_nodes, fwd_sel, bwd_sel)


def all_gather_rows(local: torch.Tensor, world: int) -> torch.Tensor:
    """Gather the ranks' padded row blocks ([n_max, ...] each) into one [world*n_max, ...] tensor (torch.distributed; the
    set-up / test path -- inside a step the rows move through ``PeerFabric``)."""
    if world == 1:
        return local
    local = local.contiguous()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)
    return out


# ------------------------------------------------------------------------------------------------------ peer fabric
class _RawCuda:
    """A raw device allocation seen through __cuda_array_interface__ (so torch can wrap it without copying)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class PeerFabric:
    """One exported buffer per rank, same layout everywhere: [flags | region 0 | region 1 | ...].

    ``signal(ch)`` (after the producer kernels, stream order) raises this step's epoch in every peer's flags;
    ``allgather(ch, parts)`` / ``reduce(ch, ...)`` / ``wait(ch)`` launch a kernel that first waits for all peers' flags of
    channel ``ch`` and then reads their memory.  Epochs grow by one per use; nothing is ever reset.  world == 1 keeps the same
    code path on a private buffer (the kernels degenerate to a wait on the rank's own flag)."""

    def __init__(self, lib, world: int, rank: int, dev: torch.device, n_channels: int, regions: Dict[str, int],
                 loopback: bool = False):
        """``loopback``: a diagnostic mode for ONE process standing in for rank ``rank`` of ``world``: every "peer" is this
        rank's own buffer (pulls copy a block onto itself, at HBM instead of NVLink speed) and a signal raises all the flags.
        The peers' blocks are never produced, so results are meaningless -- it exists to time one rank's kernels."""
        self.lib, self.world, self.rank, self.dev = lib, world, rank, dev
        self.loopback = loopback
        # row gathers: "pull" (wait for the peers' signals, read their blocks) or "push" (store this rank's block into every
        # peer's buffer, then signal); channel n_channels + ch carries the "pushes have landed" signal of gather channel ch
        import os
        self.mode = os.environ.get("B200GAT_EXCHANGE", "push")     # measured: 3 % (4 GPUs) / 1.5 % (2 GPUs) ahead of pulls
        self.n_channels = n_channels
        n_channels = 2 * n_channels
        fb = ctypes.c_size_t(0)
        lib._check(lib._lib.b200gat_peer_flag_bytes(n_channels, ctypes.byref(fb)), "peer_flag_bytes")
        self.off: Dict[str, int] = {}
        total = fb.value
        for name, nbytes in regions.items():
            self.off[name] = total
            total += _align(max(int(nbytes), 16))
        self.nbytes = total
        self.epoch = [0] * n_channels
        self._opened: List[int] = []
        self._local_ptr = None
        self.local = None
        ok, err, handle = True, "", b""
        try:
            p = ctypes.c_void_p()
            lib._check(lib._lib.b200gat_peer_alloc(self.nbytes, ctypes.byref(p)), "peer_alloc")
            self._local_ptr = p.value
            if world > 1 and not loopback:
                h = ctypes.create_string_buffer(64)
                lib._check(lib._lib.b200gat_peer_export(p, h, 64), "peer_export")
                handle = h.raw
        except Exception as exc:      # noqa: BLE001
            ok, err = False, str(exc)
        self._agree(ok, f"allocating / exporting the exchange buffer ({self.nbytes / 2**20:.0f} MiB): {err}")
        self.local = torch.as_tensor(_RawCuda(self._local_ptr, self.nbytes), device=dev)
        self.local[:fb.value].zero_()                       # flags start at 0 = "no step has signalled yet"
        ptrs = [self._local_ptr] * world
        if world > 1 and not loopback:
            gathered = [None] * world
            dist.all_gather_object(gathered, handle)
            ok, err = True, ""
            try:
                for r in range(world):
                    if r == rank:
                        continue
                    q = ctypes.c_void_p()
                    lib._check(lib._lib.b200gat_peer_open(ctypes.create_string_buffer(gathered[r], 64), ctypes.byref(q)), "peer_open")
                    self._opened.append(q.value)
                    ptrs[r] = q.value
            except Exception as exc:      # noqa: BLE001
                ok, err = False, str(exc)
            self._agree(ok, f"mapping the peers' exchange buffers (CUDA IPC): {err}")
            torch.cuda.synchronize(dev)
            dist.barrier()                                   # every rank's flags are zero before anybody signals
        self.ptrs = ptrs
        self.bases = (ctypes.c_void_p * world)(*ptrs)
        self.stats = None                                    # bench: dict name -> list of (start, end) events

    def _agree(self, ok: bool, what: str) -> None:
        """All ranks or none: a rank that failed must not leave the others blocked in the next collective."""
        flag = torch.tensor([1.0 if ok else 0.0], device=self.dev)
        if self.world > 1 and not self.loopback:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() == 0:
            self.close()
            raise RuntimeError(f"b200gat.sharded: {'failed ' + what if not ok else 'a peer rank failed ' + what.split(':')[0]}; "
                               "the row-sharded path needs peer access (CUDA IPC over NVLink) between all ranks -- there is "
                               "no host-staged fallback")

    # ---- views of the LOCAL buffer
    def view(self, name: str, shape, dtype, byte_offset: int = 0) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        o = self.off[name] + byte_offset
        assert o % 16 == 0 and o + nbytes <= self.nbytes, (name, o, nbytes, self.nbytes)
        return self.local[o:o + nbytes].view(dtype).view(tuple(int(d) for d in shape))

    def region_ptrs(self, name: str):
        """Host array of device pointers: the start of region ``name`` in every rank's buffer, as mapped on this rank."""
        return (ctypes.c_void_p * self.world)(*[self.ptrs[b] + self.off[name] for b in range(self.world)])

    def peer_ptrs(self, name: str, byte_offset: int):
        """Host array: address of (region ``name`` + byte_offset) in every PEER's buffer (self excluded), rank order."""
        others = [b for b in range(self.world) if b != self.rank]
        return (ctypes.c_void_p * max(len(others), 1))(*[self.ptrs[b] + self.off[name] + byte_offset for b in others])

    def pushed(self, ch: int) -> None:
        """After a producer kernel that stored its output into the peers' buffers itself: raise the "pushes have landed" flag
        of gather channel ``ch`` and wait for every peer's."""
        lib = self.lib
        c2 = self.n_channels + ch

        def fn():
            self.signal(c2)
            lib._check(lib._lib.b200gat_peer_wait(self.bases, self.world, self.rank, c2, self.epoch[c2], lib.stream()), "peer_wait")
        self._timed("allgather", fn)

    # ---- device-side synchronisation and transfers (all on the current stream)
    def _timed(self, what: str, fn) -> None:
        if self.stats is None:
            fn()
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        self.stats.setdefault(what, []).append((a, b))

    def signal(self, ch: int) -> None:
        self.epoch[ch] += 1
        lib = self.lib
        for r in (range(self.world) if self.loopback else (self.rank,)):
            lib._check(lib._lib.b200gat_peer_signal(self.bases, self.world, r, ch, self.epoch[ch], lib.stream()), "peer_signal")

    def wait(self, ch: int) -> None:
        lib = self.lib
        self._timed("wait", lambda: lib._check(lib._lib.b200gat_peer_wait(self.bases, self.world, self.rank, ch, self.epoch[ch],
                                                                          lib.stream()), "peer_wait"))

    def allgather(self, ch: int, parts) -> None:
        """parts: [(region name, bytes per rank block)] (1 or 2): pull block p of every part from rank p, after channel ch."""
        lib = self.lib
        offs = (ctypes.c_uint64 * len(parts))(*[self.off[n] for n, _ in parts])
        blks = (ctypes.c_uint64 * len(parts))(*[int(b) for _, b in parts])
        if self.mode == "push" and self.world > 1:
            def push():
                lib._check(lib._lib.b200gat_peer_push(self.bases, self.world, self.rank, len(parts), offs, blks, lib.stream()), "peer_push")
                self.signal(self.n_channels + ch)
                c2 = self.n_channels + ch
                lib._check(lib._lib.b200gat_peer_wait(self.bases, self.world, self.rank, c2, self.epoch[c2], lib.stream()), "peer_wait")
            self._timed("allgather", push)
            return
        self._timed("allgather", lambda: lib._check(lib._lib.b200gat_peer_allgather(
            self.bases, self.world, self.rank, ch, self.epoch[ch], len(parts), offs, blks, lib.stream()), "peer_allgather"))

    def reduce(self, ch: int, name: str, first: int, n: int, out: torch.Tensor) -> None:
        """out[i] = sum over ranks (rank order) of region[first + i] (fp32), after channel ch."""
        lib = self.lib
        assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() >= n
        self._timed("reduce", lambda: lib._check(lib._lib.b200gat_peer_reduce_f32(
            self.bases, self.world, self.rank, ch, self.epoch[ch], self.off[name], first, n, lib.ptr(out), lib.stream()), "peer_reduce"))

    def close(self) -> None:
        for q in self._opened:
            self.lib._lib.b200gat_peer_close(ctypes.c_void_p(q))
        self._opened = []
        if self._local_ptr is not None:
            self.local = None
            self.lib._lib.b200gat_peer_free(ctypes.c_void_p(self._local_ptr))
            self._local_ptr = None


# ------------------------------------------------------------------------------------------------------ trainer
class ShardedGAT:
    """Manual forward/backward of the 2-model family (custom / PyG dialect) over a row-sharded graph.

    Parameters are created with the same initialisers and in the same order as the single-GPU modules under the
    same ``torch.manual_seed``, so a sharded run starts from the identical state; ``user_emb`` rows are owned by the
    rank that owns the node, everything else is replicated and its gradient summed over the ranks.

    Why a region of the exchange buffer can be rewritten one step later without an extra barrier: a rank starts step t+1 only
    after its own step t, whose last exchange (the parameter-gradient reduce in training, the Z gather in an export) waited
    for EVERY peer's signal of that exchange -- and a peer raises that signal only after all its earlier kernels of step t,
    pulls included, have completed.  A bare ``forward()`` is followed by no such exchange, so it ends with one explicit round."""

    def __init__(self, kind: str, n_users: int, n_items: int, item_feats: torch.Tensor, edge_index: torch.Tensor,
                 hidden: int = 128, layers: int = 2, heads: int = 1, attn_dropout: float = 0.1, seed: int = 42,
                 lr: float = 1e-3, weight_decay: float = 1e-4, device: Optional[torch.device] = None,
                 feature_dtype=torch.float32, n_triples_max: int = 200_000, stream_heads: bool = False,
                 emulate: Optional[tuple] = None):
        """``emulate=(rank, world)``: diagnostic, single process: build rank ``rank``'s plan of a ``world``-rank job and run its
        kernels over a loopback fabric (see PeerFabric) -- for timing / profiling one rank's work on one GPU."""
        from . import _lib
        if feature_dtype not in (torch.float32, torch.bfloat16):
            raise NotImplementedError("feature_dtype must be float32 or bfloat16")
        if stream_heads and (kind != "pyg" or feature_dtype != torch.bfloat16):
            raise NotImplementedError("stream_heads is the bf16-tier PyG-dialect path (BASELINE config 5)")
        self.bf16 = feature_dtype == torch.bfloat16   # bf16 projection: h and the gathered dout travel (and are stored) as bf16
        from .graph import build_graph
        from .modules import CustomGAT, PyGGAT
        from .train import Adam
        self._lib = _lib
        self.kind = kind
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        if emulate is not None:
            assert self.world == 1, "emulate= is a single-process diagnostic"
            self.rank, self.world = int(emulate[0]), int(emulate[1])
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.nu, self.ni, self.n = n_users, n_items, n_users + n_items
        self.hidden, self.heads, self.n_layers = hidden, (1 if kind == "custom" else heads), layers
        self.p_drop = attn_dropout
        self.training = True
        self.policy = _lib.POLICY_CUSTOM if kind == "custom" else _lib.POLICY_PYG
        self.feat_dim = int(item_feats.shape[1])

        ei = edge_index.to(self.dev)
        self.plan = plan = make_plan(ei, n_users, n_items, self.rank, self.world)
        self.n_loc, self.n_max, self.n_pad = plan.n_loc, plan.n_max, plan.n_rows_total
        ei_p = plan.perm_map[ei]                                   # the same edges in gathered-row space
        self.g_fwd = build_graph(ei_p[:, plan.fwd_sel].contiguous(), self.n_pad)
        self.g_bwd = build_graph(ei_p[:, plan.bwd_sel].contiguous(), self.n_pad)
        # dropout masks are keyed on the ORIGINAL edge id so that the rank that owns an edge's destination (forward)
        # and the rank that owns its source (backward) regenerate the same bit
        self.perm_fwd = plan.fwd_sel[self.g_fwd.perm.long()].to(torch.int32).contiguous()
        self.perm_bwd = plan.bwd_sel[self.g_bwd.perm_csc.long()].to(torch.int32).contiguous()
        self.e_total = int(ei.shape[1])
        # local row schedules (rows are block-local ids; beg/end index the sub-graph's col / row arrays)
        self.sched_fwd = _lib.make_schedule(self.g_fwd.rowptr, plan.lo, self.n_loc, self.g_fwd.n_edges + 1)
        self.sched_bwd = _lib.make_schedule(self.g_bwd.colptr, plan.lo, self.n_loc, self.g_bwd.n_edges + 1)
        self.node_map = plan.perm_map.to(torch.int32).contiguous()
        self.node_list = plan.local_nodes.to(torch.int32).contiguous()
        del ei, ei_p
        # the forward uses the CSR side of its sub-graph only, the backward the CSC side (+ the CSR row pointers and the
        # CSR->CSC map for the per-destination sums): drop the rest (8 x 4 B per edge of config 5's 800 M edges)
        self.g_fwd.row = self.g_fwd.perm_csc = self.g_fwd.csr2csc = self.g_fwd.perm = None
        self.g_bwd.col = self.g_bwd.perm = self.g_bwd.perm_csc = None
        self.g_fwd.sched_fwd = self.g_fwd.sched_bwd = self.g_bwd.sched_fwd = self.g_bwd.sched_bwd = None

        torch.manual_seed(seed)
        full = (CustomGAT(n_users, n_items, self.feat_dim, hidden, layers) if kind == "custom"
                else PyGGAT(n_users, n_items, self.feat_dim, hidden, layers, heads, attn_dropout))
        P = torch.nn.Parameter
        self.user_emb = P(full.user_emb.weight.detach()[self.rank::self.world].clone().to(self.dev))
        self.item_proj = full.item_proj.to(self.dev)
        self.feats_loc = item_feats[self.rank::self.world].to(self.dev).contiguous()
        self.W, self.a_src, self.a_dst, self.bias = [], [], [], []
        for l in range(layers):
            lay = full.layers[l] if kind == "custom" else full.convs[l]
            self.W.append(P(lay.lin.weight.detach().clone().to(self.dev)))
            a_s, a_d = (lay.a_src, lay.a_dst) if kind == "custom" else (lay.att_src, lay.att_dst)
            self.a_src.append(P(a_s.detach().clone().to(self.dev).view(self.heads, hidden)))
            self.a_dst.append(P(a_d.detach().clone().to(self.dev).view(self.heads, hidden)))
            self.bias.append(None if kind == "custom" else P(lay.bias.detach().clone().to(self.dev)))
        del full
        self.replicated = list(self.item_proj.parameters()) + self.W + self.a_src + self.a_dst + [b for b in self.bias if b is not None]
        self.n_rep = sum(p.numel() for p in self.replicated)
        self.opt = Adam([self.user_emb] + self.replicated, lr=lr, weight_decay=weight_decay)   # the reference's update rule
        self.step_no = 0
        self.seed = seed

        # ---- exchange buffer: regions and channels
        H, C, L, n_max, n_pad = self.heads, hidden, layers, self.n_max, self.n_pad
        self.x_exchange = H > 1                  # heads > 1: move the layer input (F_in wide), project all rows on every rank
        self.stream = bool(stream_heads)
        hsz = 2 if self.bf16 else 4
        self.s_max = int(n_triples_max)
        regions: Dict[str, int] = {}
        if self.stream:
            # per-head streaming (config 5): two alternating bf16 regions for the gathered layer inputs (the idle one holds the
            # gathered dout in the backward), one head's nodestat and per-destination partial sums at a time
            regions["X0"] = n_pad * C * 2
            regions["X1"] = n_pad * C * 2
            regions["NSh"] = n_pad * 16
            regions["DSDh"] = n_pad * 4
        for l in range(L if not self.stream else 0):
            regions[f"F{l}"] = n_pad * (C * 4 if self.x_exchange else H * C * hsz)
            if not self.x_exchange:
                regions[f"S{l}"] = n_pad * 2 * H * 4
            regions[f"D{l}"] = n_pad * C * (4 if l == 0 else hsz)     # D0 doubles as the export's fp32 gather buffer
            regions[f"NS{l}"] = n_pad * H * 16
            regions[f"DSD{l}"] = n_pad * H * 4
        regions["Z"] = n_max * C * 4
        regions["COEF"] = 2 * self.s_max * 4
        regions["LOSS"] = 16
        regions["G"] = self.n_rep * 4
        self.CH_F = lambda l: l
        self.CH_Z, self.CH_C = L, L + 1
        self.CH_B = lambda l: L + 2 + l
        self.CH_R = lambda l: 2 * L + 2 + l
        self.CH_G = 3 * L + 2
        self.CH_HA, self.CH_HR1, self.CH_HB, self.CH_HR2 = 3 * L + 3, 3 * L + 4, 3 * L + 5, 3 * L + 6   # per-head rounds (streaming)
        self.fab = PeerFabric(_lib, self.world, self.rank, self.dev, 3 * L + 7, regions, loopback=emulate is not None)
        # fused projection + exchange: fp32 tier, heads == 1, 128-wide layers, tensor-core GEMMs, push mode, real peers
        self.fused_push = (self.fab.mode == "push" and self.world > 1 and emulate is None and not self.bf16 and H == 1 and C == 128
                           and _lib.get_gemm_mode() == _lib.GEMM_TF32X3 and os.environ.get("B200GAT_FUSED_PUSH", "1") != "0")
        self.comm_bytes_per_step = 0            # bytes this rank pulls from its peers per training step
        self.comm_now = 0
        self._loss_ws = None
        self.saved = []

    def close(self) -> None:
        self.fab.close()

    # -------------------------------------------------------------------------------------------- helpers
    def _empty(self, *shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.dev)

    def _layer_seed(self, layer: int) -> int:
        return (self.seed * 1_000_003 + self.step_no * 101 + layer) & (2 ** 62 - 1)

    def _block(self, name: str, width: int, dtype) -> torch.Tensor:
        """This rank's [n_max, width] block of a gathered region (what the producer kernels write)."""
        esz = torch.empty((), dtype=dtype).element_size()
        return self.fab.view(name, (self.n_max, width), dtype, self.rank * self.n_max * width * esz)

    def _full(self, name: str, width: int, dtype) -> torch.Tensor:
        return self.fab.view(name, (self.n_pad, width), dtype)

    def _pulled(self, parts) -> None:
        self.comm_now += sum(b for _, b in parts) * (self.world - 1)

    # -------------------------------------------------------------------------------------------- forward
    def forward(self, _exchange_follows: bool = False) -> torch.Tensor:
        """Returns the local rows of Z ([n_max, C], the first n_loc are real; lives in the exchange buffer); keeps what the
        backward needs in ``self.saved``."""
        if self.stream:
            return self._forward_stream(_exchange_follows)
        lib, H, C, fab = self._lib, self.heads, self.hidden, self.fab
        st = lib.stream()
        L = self.n_layers
        self.comm_now = 0
        h_dt = torch.bfloat16 if self.bf16 else torch.float32
        hsz = 2 if self.bf16 else 4
        # layer-0 input: [user rows | item_proj(features)] without a concat copy (node_features, train_gat_custom.py:105-109)
        x = self._block("F0", C, torch.float32) if self.x_exchange else self._empty(self.n_max, C)
        cu = self.plan.cu
        x[:cu].copy_(self.user_emb.detach())
        dwb = lib.dense_workspace_bytes(H, C, max(C, self.feat_dim))
        dws = self._empty(dwb, dtype=torch.uint8)
        if self.plan.ci:
            lib.call("b200gat_linear_tc_f32" if self.bf16 else "b200gat_linear_f32", lib.ptr(self.feats_loc), lib.ptr(self.item_proj.weight),
                     lib.ptr(self.item_proj.bias), self.plan.ci, self.feat_dim, C, lib.ptr(x, cu * C), C, lib.ptr(dws), dwb, st)
        self.saved = []
        p = self.p_drop if self.training else 0.0
        for l in range(L):
            last = l == L - 1
            if self.x_exchange:
                # x rows travel; every rank projects all rows (padding rows of the peers' blocks are never referenced by an edge)
                fab.signal(self.CH_F(l))
                parts = [(f"F{l}", self.n_max * C * 4)]
                fab.allgather(self.CH_F(l), parts)
                self._pulled(parts)
                x_full = self._full(f"F{l}", C, torch.float32)
                h_full = self._empty(self.n_pad, H * C, dtype=h_dt)
                s_full = self._empty(self.n_pad, 2 * H)
                lib.call("b200gat_project_bf16" if self.bf16 else "b200gat_project_f32", lib.ptr(x_full), lib.ptr(self.W[l]),
                         lib.ptr(self.a_src[l]), lib.ptr(self.a_dst[l]), self.n_pad, C, H, C, lib.ptr(h_full), lib.ptr(s_full),
                         lib.ptr(dws), dwb, st)
            else:
                h_loc, s_loc = self._block(f"F{l}", H * C, h_dt), self._block(f"S{l}", 2 * H, torch.float32)
                parts = [(f"F{l}", self.n_max * H * C * hsz), (f"S{l}", self.n_max * 2 * H * 4)]
                if self.fused_push:
                    # the GEMM's epilogue stores h tiles and logits into the peers' buffers as well (csrc/gemm_tc.cu)
                    lib.call("b200gat_project_push_f32", lib.ptr(x), lib.ptr(self.W[l]), lib.ptr(self.a_src[l]), lib.ptr(self.a_dst[l]),
                             self.n_loc, C, H, C, lib.ptr(h_loc), lib.ptr(s_loc), fab.peer_ptrs(f"F{l}", self.rank * parts[0][1]),
                             fab.peer_ptrs(f"S{l}", self.rank * parts[1][1]), self.world - 1, lib.ptr(dws), dwb, st)
                    fab.pushed(self.CH_F(l))
                else:
                    lib.call("b200gat_project_bf16" if self.bf16 else "b200gat_project_f32", lib.ptr(x), lib.ptr(self.W[l]),
                             lib.ptr(self.a_src[l]), lib.ptr(self.a_dst[l]), self.n_loc, C, H, C, lib.ptr(h_loc), lib.ptr(s_loc),
                             lib.ptr(dws), dwb, st)
                    fab.signal(self.CH_F(l))
                    fab.allgather(self.CH_F(l), parts)
                self._pulled(parts)
                h_full, s_full = self._full(f"F{l}", H * C, h_dt), self._full(f"S{l}", 2 * H, torch.float32)
            if last:
                out = fab.view("Z", (self.n_max, C), torch.float32)
            elif self.x_exchange:
                out = self._block(f"F{l + 1}", C, torch.float32)           # the next layer's input, produced in place
            else:
                out = self._empty(self.n_max, C)
            rowstat = self._empty(self.n_loc, H, 2)
            out_heads = self._empty(self.n_loc, H, C) if H > 1 else None
            seed = self._layer_seed(l)
            sf = self.sched_fwd
            lib.call("b200gat_edge_fwd_bf16" if self.bf16 else "b200gat_edge_fwd_f32", lib.ptr(h_full), lib.ptr(s_full),
                     lib.ptr(sf.sched), sf.n_sched, lib.ptr(sf.table),
                     sf.n_long, lib.ptr(sf.partial(H * (C + 4))), lib.ptr(self.g_fwd.col), lib.ptr(self.perm_fwd), self.plan.lo,
                     H, C, self.policy, 0.2, lib.ptr(self.bias[l]), lib.ptr(out), lib.ptr(out_heads), lib.ptr(rowstat), p, seed, st)
            self.saved.append((x, h_full, s_full, rowstat, out if H == 1 else out_heads, p, seed))
            x = out
        if not _exchange_follows and self.world > 1:
            fab.signal(self.CH_Z)              # a bare forward: one round so that the next call may rewrite the regions
            fab.wait(self.CH_Z)
        return x

    # -------------------------------------------------------------------------------------------- per-head streaming
    def _x0_bf16(self, dst: torch.Tensor) -> None:
        """Layer-0 input rows [user rows | item_proj(features)] as bf16 into ``dst`` [n_max, C] (recomputed in the backward
        instead of being kept: it is a function of parameters and inputs only).  The item rows go through an fp32 staging
        buffer of at most 2 M rows at a time (config 5: 15 GB if done in one piece)."""
        lib, C = self._lib, self.hidden
        st = lib.stream()
        cu, ci = self.plan.cu, self.plan.ci
        if cu:
            lib.call("b200gat_cast_bf16", lib.ptr(self.user_emb.detach()), lib.ptr(dst), cu * C, 1.0, st)
        dwb = lib.dense_workspace_bytes(1, C, max(C, self.feat_dim))
        dws = self._empty(dwb, dtype=torch.uint8)
        chunk = 2_000_000
        x = self._empty(min(chunk, max(ci, 1)), C)
        for lo in range(0, ci, chunk):
            n = min(chunk, ci - lo)
            lib.call("b200gat_linear_tc_f32", lib.ptr(self.feats_loc, lo * self.feat_dim), lib.ptr(self.item_proj.weight),
                     lib.ptr(self.item_proj.bias), n, self.feat_dim, C, lib.ptr(x), C, lib.ptr(dws), dwb, st)
            lib.call("b200gat_cast_bf16", lib.ptr(x), lib.ptr(dst, (cu + lo) * C), n * C, 1.0, st)
        if self.n_max > self.n_loc:
            dst[self.n_loc:].zero_()              # padding rows travel with the block: keep them finite

    def _forward_stream(self, _exchange_follows: bool) -> torch.Tensor:
        """heads > 1 one head at a time: [N, heads*C] never exists.  Per layer: the bf16 input rows are exchanged once, then for
        every head the projection of ALL rows (h_head [N, C] bf16, reused buffer) and the fused edge forward of the local rows,
        which accumulates the head mean into ``out``.  Saved for the backward: the layer input rows (bf16, layers > 0), and per
        head the per-node scalars s and rowstat."""
        lib, H, C, fab = self._lib, self.heads, self.hidden, self.fab
        st = lib.stream()
        L = self.n_layers
        self.comm_now = 0
        blk = self.n_max * C * 2
        xb = fab.view("X0", (self.n_max, C), torch.bfloat16, self.rank * blk)
        self._x0_bf16(xb)
        self.saved = []
        p = self.p_drop if self.training else 0.0
        dwb = lib.dense_workspace_bytes(1, C, C)
        dws = self._empty(dwb, dtype=torch.uint8)
        h_head = self._empty(self.n_pad, C, dtype=torch.bfloat16)
        sf = self.sched_fwd
        for l in range(L):
            reg = f"X{l % 2}"
            fab.signal(self.CH_F(l))
            parts = [(reg, blk)]
            fab.allgather(self.CH_F(l), parts)
            self._pulled(parts)
            x_full = fab.view(reg, (self.n_pad, C), torch.bfloat16)
            x_keep = None if l == 0 else fab.view(reg, (self.n_max, C), torch.bfloat16, self.rank * blk).clone()
            last = l == L - 1
            out = fab.view("Z", (self.n_max, C), torch.float32) if last else self._empty(self.n_max, C)
            seed = self._layer_seed(l)
            s_heads, rs_heads = [], []
            for hh in range(H):
                s_h = self._empty(self.n_pad, 2)
                rs_h = self._empty(self.n_loc, 2)
                lib.call("b200gat_project_bf16_ex", lib.ptr(x_full), 1, lib.ptr(self.W[l], hh * C * C), lib.ptr(self.a_src[l], hh * C),
                         lib.ptr(self.a_dst[l], hh * C), self.n_pad, C, 1, C, lib.ptr(h_head), lib.ptr(s_h), lib.ptr(dws), dwb, st)
                lib.call("b200gat_edge_fwd_stream_bf16", lib.ptr(h_head), lib.ptr(s_h), lib.ptr(sf.sched), sf.n_sched, lib.ptr(sf.table),
                         sf.n_long, lib.ptr(sf.partial(C + 4)), lib.ptr(self.g_fwd.col), lib.ptr(self.perm_fwd), self.plan.lo, C,
                         self.policy, 0.2, lib.ptr(self.bias[l]) if hh == 0 else None, lib.ptr(out), lib.ptr(rs_h), p, seed + 7919 * hh,
                         1.0 / H, int(hh > 0), st)
                s_heads.append(s_h)
                rs_heads.append(rs_h)
            self.saved.append((x_keep, s_heads, rs_heads, p, seed))
            if not last:        # the next layer's input rows, bf16, straight into the other region's block
                nxt = fab.view(f"X{(l + 1) % 2}", (self.n_max, C), torch.bfloat16, self.rank * blk)
                lib.call("b200gat_cast_bf16", lib.ptr(out), lib.ptr(nxt), self.n_max * C, 1.0, st)
        del h_head
        if not _exchange_follows and self.world > 1:
            fab.signal(self.CH_Z)
            fab.wait(self.CH_Z)
        return out

    def _backward_stream(self, dout_box: list, grads: dict) -> torch.Tensor:
        """Backward of the streamed layers.  Nothing per-head was kept but scalars: h of the local source rows is re-projected
        per head, and t_i = sum_k alpha_ik dalpha_ik comes out of a first pass over the edges (two-phase backward, edge.cu)."""
        lib, H, C, fab = self._lib, self.heads, self.hidden, self.fab
        st = lib.stream()
        L, lo, P_ = self.n_layers, self.plan.lo, self.world
        blk = self.n_max * C * 2
        dwb = lib.dense_workspace_bytes(1, C, C)
        dws = self._empty(dwb, dtype=torch.uint8)
        sb = self.sched_bwd
        e_loc = max(self.g_bwd.n_edges, 1)
        dout = dout_box.pop()                 # the only reference: it is released below, before dx / dh are allocated
        for l in reversed(range(L)):
            x_keep, s_heads, rs_heads, p, seed = self.saved[l]
            if x_keep is None:
                x_keep = self._empty(self.n_max, C, dtype=torch.bfloat16)
                self._x0_bf16(x_keep)
            db = torch.empty_like(self.bias[l])
            lib.call("b200gat_colsum_f32", lib.ptr(dout), self.n_loc, C, lib.ptr(db), lib.ptr(dws), dwb, st)
            grads[self.bias[l]] = db
            # gathered dout, bf16, already carrying the 1/heads of the head mean (exact: a power of two)
            dg = fab.view("X0", (self.n_max, C), torch.bfloat16, self.rank * blk)
            lib.call("b200gat_cast_bf16", lib.ptr(dout), lib.ptr(dg), self.n_max * C, 1.0 / H, st)
            dout = None                           # [n_max, C] fp32 (15 GB at config 5): free it before the next allocations
            fab.signal(self.CH_B(l))
            parts = [("X0", blk)]
            fab.allgather(self.CH_B(l), parts)
            self._pulled(parts)
            dout_full = fab.view("X0", (self.n_pad, C), torch.bfloat16)
            dx = self._empty(self.n_max, C)
            dW = torch.empty_like(self.W[l])
            da_s, da_d = torch.empty_like(self.a_src[l]), torch.empty_like(self.a_dst[l])
            h_loc = self._empty(self.n_loc, C, dtype=torch.bfloat16)
            s_tmp = self._empty(self.n_loc, 2)
            dh = self._empty(self.n_loc, C)
            de = self._empty(e_loc)
            ds = self._empty(self.n_loc, 2)
            red = self._empty(max(self.n_loc, 1))
            ns_blk = fab.view("NSh", (self.n_max, 4), torch.float32, self.rank * self.n_max * 16)
            ns_full = fab.view("NSh", (self.n_pad, 4), torch.float32)
            part = fab.view("DSDh", (self.n_pad,), torch.float32)
            ns_parts = [("NSh", self.n_max * 16)]
            for hh in range(H):
                s_h, rs_h = s_heads[hh], rs_heads[hh]
                W_h, as_h, ad_h = lib.ptr(self.W[l], hh * C * C), lib.ptr(self.a_src[l], hh * C), lib.ptr(self.a_dst[l], hh * C)
                lib.call("b200gat_project_bf16_ex", lib.ptr(x_keep), 1, W_h, as_h, ad_h, self.n_loc, C, 1, C, lib.ptr(h_loc), lib.ptr(s_tmp),
                         lib.ptr(dws), dwb, st)
                lib.call("b200gat_node_stat_f32", lib.ptr(s_h), lib.ptr(rs_h), self.n_loc, lo, 1, lib.ptr(ns_blk), st)
                fab.signal(self.CH_HA)
                fab.allgather(self.CH_HA, ns_parts)
                self._pulled(ns_parts)
                # the kernels index h by gathered row (row_offset + r): hand them the local block shifted back by row_offset rows
                lib.call("b200gat_edge_bwd_phase1_bf16", lib.ptr(h_loc, -lo * C), lib.ptr(s_h), lib.ptr(dout_full), lib.ptr(ns_full),
                         lib.ptr(sb.sched), sb.n_sched, lib.ptr(sb.table), sb.n_long, lib.ptr(sb.partial(C + 4)), lib.ptr(self.g_bwd.row),
                         lib.ptr(self.perm_bwd), lo, C, self.policy, 0.2, lib.ptr(dh), lib.ptr(de), lib.ptr(ds), 2, p, seed + 7919 * hh, st)
                lib.call("b200gat_ds_dst_f32", lib.ptr(de), lib.ptr(self.g_bwd.rowptr), lib.ptr(self.g_bwd.csr2csc), self.n_pad,
                         self.g_bwd.n_edges, 1, lib.ptr(part), 1, st)
                fab.signal(self.CH_HR1)
                fab.reduce(self.CH_HR1, "DSDh", lo, self.n_loc, red)                   # t of the local destinations
                lib.call("b200gat_node_stat_set_t_f32", lib.ptr(ns_blk), lib.ptr(red), self.n_loc, st)
                fab.signal(self.CH_HB)
                fab.allgather(self.CH_HB, ns_parts)
                self._pulled(ns_parts)
                lib.call("b200gat_edge_bwd_phase2_f32", lib.ptr(de), lib.ptr(self.g_bwd.colptr), lib.ptr(self.g_bwd.row), lib.ptr(s_h),
                         lib.ptr(ns_full), self.n_loc, lo, self.policy, 0.2, lib.ptr(ds), 2, st)
                lib.call("b200gat_ds_dst_f32", lib.ptr(de), lib.ptr(self.g_bwd.rowptr), lib.ptr(self.g_bwd.csr2csc), self.n_pad,
                         self.g_bwd.n_edges, 1, lib.ptr(part), 1, st)
                fab.signal(self.CH_HR2)
                fab.reduce(self.CH_HR2, "DSDh", lo, self.n_loc, red)
                self.comm_now += 2 * self.n_loc * 4 * (P_ - 1)
                ds[:, 1] = red[:self.n_loc]
                lib.call("b200gat_project_bwd_bf16_ex", lib.ptr(x_keep), 1, W_h, as_h, ad_h, lib.ptr(dh), lib.ptr(ds), self.n_loc, C, 1, C,
                         lib.ptr(dx), int(hh > 0), lib.ptr(dW, hh * C * C), lib.ptr(da_s, hh * C), lib.ptr(da_d, hh * C),
                         lib.ptr(dws), dwb, st)
            grads[self.W[l]], grads[self.a_src[l]], grads[self.a_dst[l]] = dW, da_s, da_d
            dout = dx
            self.saved[l] = None
        return dout

    # -------------------------------------------------------------------------------------------- loss + backward
    def loss_and_backward(self, z_loc: torch.Tensor, u, i, j, loss_kind: str = "bpr") -> torch.Tensor:
        lib, H, C, fab = self._lib, self.heads, self.hidden, self.fab
        st = lib.stream()
        L, P_, r = self.n_layers, self.world, self.rank
        hsz = 2 if self.bf16 else 4
        assert z_loc.data_ptr() == fab.view("Z", (1,), torch.float32).data_ptr(), "z_loc must be the tensor forward() returned"
        S = int(u.shape[0])
        if S > self.s_max:
            raise RuntimeError(f"{S} triples > n_triples_max={self.s_max} given at construction")
        kind = lib.LOSS_BPR if loss_kind == "bpr" else lib.LOSS_BCE
        ws_bytes = lib.loss_workspace_bytes(self.n, S)
        if self._loss_ws is None or self._loss_ws.numel() < ws_bytes:
            self._loss_ws = self._empty(ws_bytes, dtype=torch.uint8)
        ws = self._loss_ws
        # ---- loss forward on this rank's slice of the triples; rows come straight out of the owners' Z blocks
        per = (S + P_ - 1) // P_
        t0, t1 = min(r * per, S), min((r + 1) * per, S)
        coef_loc = fab.view("COEF", (2 * S,), torch.float32)
        coef_loc.zero_()                                      # the reduce below sums the ranks' disjoint slices
        loss_part = fab.view("LOSS", (1,), torch.float32)
        z_blocks = fab.region_ptrs("Z")
        fab.signal(self.CH_Z)
        fab.wait(self.CH_Z)                                   # every rank's Z block is complete
        lib.call("b200gat_rank_loss_fwd_peer_f32", z_blocks, P_, self.n_max, self.nu, self.ni, C, lib.ptr(u), lib.ptr(i), lib.ptr(j),
                 S, t0, t1 - t0, lib.ptr(self.node_map), kind, 1, lib.ptr(coef_loc), lib.ptr(loss_part), lib.ptr(ws), ws_bytes, st)
        self.comm_now += 3 * (t1 - t0) * C * 4 * (P_ - 1) // P_
        fab.signal(self.CH_C)
        coef = self._empty(2 * S)
        loss = self._empty(1)
        fab.reduce(self.CH_C, "COEF", 0, 2 * S, coef)
        fab.reduce(self.CH_C, "LOSS", 0, 1, loss)
        self.comm_now += 2 * S * 4 * (P_ - 1)
        # ---- gradient rows of this rank's nodes (partner rows again read from the owners), straight into the exchange block
        one = torch.ones(1, dtype=torch.float32, device=self.dev)
        dout = self._block(f"D{L - 1}", C, torch.float32) if not self.bf16 else self._empty(self.n_max, C)
        if self.stream:
            dout.zero_()                         # rows [n_loc, n_max) are cast and exchanged with the rest: keep them finite
        lib.call("b200gat_rank_loss_bwd_peer_f32", z_blocks, P_, self.n_max, self.nu, self.ni, C, lib.ptr(u), lib.ptr(i), lib.ptr(j), S,
                 lib.ptr(self.node_map), kind, lib.ptr(coef), lib.ptr(one), lib.ptr(self.node_list), self.n_loc, lib.ptr(dout), None,
                 lib.ptr(ws), ws_bytes, st)
        self.comm_now += 4 * S * C * 4 * (P_ - 1) // (P_ * P_)
        grads = {}
        dwb = lib.dense_workspace_bytes(H, C, max(C, self.feat_dim))
        dws = self._empty(dwb, dtype=torch.uint8)
        d_dt = torch.bfloat16 if self.bf16 else torch.float32
        if self.stream:
            box = [dout]
            del dout
            dout = self._backward_stream(box, grads)
        for l in (reversed(range(L)) if not self.stream else ()):
            x, h_full, s_full, rowstat, out_h, p, seed = self.saved[l]
            nodestat = self._block(f"NS{l}", H * 4, torch.float32)
            db = torch.empty_like(self.bias[l]) if self.bias[l] is not None else None
            dout_g = self._block(f"D{l}", C, torch.bfloat16) if self.bf16 else None
            if not self.bf16:
                assert dout.data_ptr() == self._block(f"D{l}", C, torch.float32).data_ptr()
            lib.call("b200gat_node_prep_f32", lib.ptr(dout), lib.ptr(out_h), lib.ptr(self.bias[l] if H == 1 else None),
                     lib.ptr(s_full), lib.ptr(rowstat), self.n_loc, self.plan.lo, H, C, lib.ptr(nodestat), lib.ptr(db),
                     lib.ptr(dout_g), lib.ptr(dws), dwb, st)
            fab.signal(self.CH_B(l))
            parts = [(f"D{l}", self.n_max * C * hsz), (f"NS{l}", self.n_max * H * 16)]
            # fused mode, l < L-1: this layer's dout block already sits in the peers' buffers (stored by the dx GEMM of layer l+1)
            fab.allgather(self.CH_B(l), parts[1:] if (self.fused_push and l < L - 1) else parts)
            self._pulled(parts)
            dout_full, nodestat_full = self._full(f"D{l}", C, d_dt), self._full(f"NS{l}", H * 4, torch.float32)
            dh = self._empty(self.n_loc, H * C)
            de = self._empty(max(self.g_bwd.n_edges, 1), H)
            ds = self._empty(self.n_loc, 2 * H)
            sb = self.sched_bwd
            lib.call("b200gat_edge_bwd_bf16" if self.bf16 else "b200gat_edge_bwd_f32", lib.ptr(h_full), lib.ptr(s_full),
                     lib.ptr(dout_full), lib.ptr(nodestat_full),
                     lib.ptr(sb.sched), sb.n_sched, lib.ptr(sb.table), sb.n_long, lib.ptr(sb.partial(H * C + 4)),
                     lib.ptr(self.g_bwd.row), lib.ptr(self.perm_bwd), self.plan.lo, H, C, self.policy, 0.2, lib.ptr(dh),
                     lib.ptr(de), lib.ptr(ds), 2 * H, p, seed, st)
            ds_part = self._full(f"DSD{l}", H, torch.float32)               # partial sums over this rank's edges, all rows
            lib.call("b200gat_ds_dst_f32", lib.ptr(de), lib.ptr(self.g_bwd.rowptr), lib.ptr(self.g_bwd.csr2csc), self.n_pad,
                     self.g_bwd.n_edges, H, lib.ptr(ds_part), H, st)
            fab.signal(self.CH_R(l))
            ds_dst = self._empty(max(self.n_loc, 1), H)
            fab.reduce(self.CH_R(l), f"DSD{l}", self.plan.lo * H, self.n_loc * H, ds_dst)
            self.comm_now += self.n_loc * H * 4 * (P_ - 1)
            ds[:, H:] = ds_dst[:self.n_loc]
            del de
            # dx of layer l is the dout of layer l-1: in the fp32 tier it is produced straight into that layer's exchange block
            dx = self._block(f"D{l - 1}", C, torch.float32) if (l >= 1 and not self.bf16) else self._empty(self.n_max, C)
            dW, da_s, da_d = torch.empty_like(self.W[l]), torch.empty_like(self.a_src[l]), torch.empty_like(self.a_dst[l])
            if self.fused_push and l >= 1:
                lib.call("b200gat_project_bwd_push_f32", lib.ptr(x), lib.ptr(self.W[l]), lib.ptr(self.a_src[l]), lib.ptr(self.a_dst[l]),
                         lib.ptr(dh), lib.ptr(ds), self.n_loc, C, H, C, lib.ptr(dx),
                         fab.peer_ptrs(f"D{l - 1}", self.rank * self.n_max * C * 4), self.world - 1, lib.ptr(dW), lib.ptr(da_s),
                         lib.ptr(da_d), lib.ptr(dws), dwb, st)
            else:
                lib.call("b200gat_project_bwd_bf16" if self.bf16 else "b200gat_project_bwd_f32", lib.ptr(x), lib.ptr(self.W[l]),
                         lib.ptr(self.a_src[l]), lib.ptr(self.a_dst[l]), lib.ptr(dh), lib.ptr(ds), self.n_loc, C, H, C, lib.ptr(dx),
                         lib.ptr(dW), lib.ptr(da_s), lib.ptr(da_d), lib.ptr(dws), dwb, st)
            grads[self.W[l]], grads[self.a_src[l]], grads[self.a_dst[l]] = dW, da_s, da_d
            if db is not None:
                grads[self.bias[l]] = db
            dout = dx
        # input features (node_features backward): user rows take their gradient rows, item_proj its weight / bias gradients
        cu, ci = self.plan.cu, self.plan.ci
        self.user_emb.grad = dout[:cu].clone()
        dpw, dpb = torch.empty_like(self.item_proj.weight), torch.empty_like(self.item_proj.bias)
        lib.call("b200gat_linear_bwd_f32", lib.ptr(self.feats_loc), lib.ptr(dout, cu * C), C, ci, self.feat_dim, C, lib.ptr(dpw),
                 lib.ptr(dpb), lib.ptr(dws), dwb, st)
        grads[self.item_proj.weight], grads[self.item_proj.bias] = dpw, dpb
        # replicated parameters: sum of the ranks' gradients, in rank order on every rank (bitwise identical everywhere)
        g_loc = fab.view("G", (self.n_rep,), torch.float32)
        torch.cat([grads[p_].reshape(-1) for p_ in self.replicated], out=g_loc)
        fab.signal(self.CH_G)
        flat = self._empty(self.n_rep)
        fab.reduce(self.CH_G, "G", 0, self.n_rep, flat)
        self.comm_now += self.n_rep * 4 * (P_ - 1)
        off = 0
        for p_ in self.replicated:
            k = p_.numel()
            p_.grad = flat[off:off + k].view_as(p_)
            off += k
        self.saved = []
        self.comm_bytes_per_step = self.comm_now
        return loss.view(())

    def train_step(self, u, i, j, loss_kind: str = "bpr") -> torch.Tensor:
        for p_ in [self.user_emb] + self.replicated:      # last step's gradients (config 5: 5 GB of user rows) are dead weight now
            p_.grad = None
        z = self.forward(_exchange_follows=True)
        loss = self.loss_and_backward(z, u, i, j, loss_kind)
        self.opt.step()
        self.step_no += 1
        return loss

    @torch.no_grad()
    def export_item_embeddings(self) -> torch.Tensor:
        """Config 4: forward-only, returns Z[n_users:] gathered on every rank (tools/export_item_embeddings.py:140-142)."""
        was, self.training = self.training, False
        z_loc = self.forward(_exchange_follows=True)
        self.training = was
        self.saved = []
        fab, C = self.fab, self.hidden
        # the gather of the ranks' Z blocks doubles as the end-of-call round that lets the next call rewrite the regions
        zg = fab.view("D0", (self.n_pad, C), torch.float32)
        zg[self.plan.lo:self.plan.lo + self.n_max].copy_(z_loc)
        fab.signal(self.CH_B(0))
        parts = [("D0", self.n_max * C * 4)]
        fab.allgather(self.CH_B(0), parts)
        self._pulled(parts)
        self.comm_bytes_per_step = self.comm_now
        return zg[self.plan.perm_map[self.nu:]]                   # back to node order, item rows only


# ------------------------------------------------------------------------------------------------------ bench (N > 1)
def _parity_vs_single(tr: "ShardedGAT", cfg, feats, ei, triples, rank: int, world: int, dev) -> Optional[dict]:
    """One eval-mode forward + loss + backward on the sharded trainer against the single-GPU module path holding the SAME
    parameters (after the timed steps), every rank checking its own rows: max relative differences (on the tensor's scale)."""
    import b200gat
    u, i, j = triples
    nu, ni = tr.nu, tr.ni
    tr.training = False
    z_loc = tr.forward(_exchange_follows=True)
    loss = tr.loss_and_backward(z_loc, u, i, j, cfg["loss"])
    z_rows = z_loc[:tr.n_loc].clone()
    ue = torch.zeros(nu, tr.hidden, device=dev)
    ue[rank::world] = tr.user_emb.detach()
    if world > 1:
        dist.all_reduce(ue)
    fdt = torch.bfloat16 if tr.bf16 else torch.float32
    custom = tr.kind == "custom"
    m = (b200gat.CustomGAT(nu, ni, tr.feat_dim, tr.hidden, tr.n_layers, feature_dtype=fdt) if custom
         else b200gat.PyGGAT(nu, ni, tr.feat_dim, tr.hidden, tr.n_layers, tr.heads, 0.1, feature_dtype=fdt)).to(dev).eval()
    lays = m.layers if custom else m.convs
    with torch.no_grad():
        m.user_emb.weight.copy_(ue)
        m.item_proj.weight.copy_(tr.item_proj.weight)
        m.item_proj.bias.copy_(tr.item_proj.bias)
        for l, lay in enumerate(lays):
            lay.lin.weight.copy_(tr.W[l])
            a_s, a_d = (lay.a_src, lay.a_dst) if custom else (lay.att_src, lay.att_dst)
            a_s.copy_(tr.a_src[l].view_as(a_s))
            a_d.copy_(tr.a_dst[l].view_as(a_d))
            if not custom:
                lay.bias.copy_(tr.bias[l])
    z = m(feats.to(dev), ei.to(dev))
    ref_loss = (b200gat.bpr_loss if cfg["loss"] == "bpr" else b200gat.bce_loss)(z, nu, u, i, j)
    ref_loss.backward()
    rel = lambda a, b: float((a.detach().double() - b.detach().double()).abs().max() / b.detach().double().abs().max().clamp_min(1e-30))
    res = torch.tensor([rel(z_rows, z[tr.plan.local_nodes]), rel(loss, ref_loss), rel(tr.W[0].grad, lays[0].lin.weight.grad),
                        rel(tr.W[-1].grad, lays[-1].lin.weight.grad),
                        rel(tr.user_emb.grad, m.user_emb.weight.grad[rank::world]) if tr.plan.cu else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(res, op=dist.ReduceOp.MAX)
    tr.training = True
    del m, z
    torch.cuda.empty_cache()
    names = ["z_rows", "loss", "dW_layer0", "dW_last", "d_user_emb_rows"]
    return {"mode": "eval (no dropout), same parameters, after the timed steps; max over ranks of max|a-b| / max|b|",
            **{k_: float(v) for k_, v in zip(names, res.tolist())}}


def bench_main(args, cfg, rank: int, world: int, dev: torch.device) -> None:
    """bench.py's N>1 leg: strong scaling of the same workload, max-over-ranks device time."""
    from . import _lib, synth
    import bench as B
    import sys
    t_start = time.time()

    def note(what: str) -> None:          # progress on stderr (rank 0): a full-size config-5 run spends minutes in set-up
        if rank == 0:
            torch.cuda.synchronize()
            print(f"[bench +{time.time() - t_start:6.1f}s] {what}; HBM allocated {torch.cuda.memory_allocated(dev) / 2**30:.1f} GiB "
                  f"(peak {torch.cuda.max_memory_allocated(dev) / 2**30:.1f})", file=sys.stderr, flush=True)

    nu, ni, n_inter, k = B.graph_dims(cfg, synth)
    bf16 = cfg["tier"] == "bf16"
    export = cfg["mode"] == "export"
    L, H, C = cfg["layers"], cfg["heads"], cfg["hidden"]
    e = 2 * n_inter + k * ni
    on_device = e >= 50_000_000        # the host generator needs minutes and ~100 B/edge of host memory PER RANK at this size
    if on_device:
        # rank 0 draws the edge list and broadcasts it (the device generator is not bitwise reproducible between processes);
        # the item features are element-wise Philox draws, which every rank can repeat for itself
        if rank == 0:
            ei, _ = synth.make_graph_device(nu, ni, n_inter, k, dev)
        else:
            ei = torch.empty((2, e), dtype=torch.int64, device=dev)
        if world > 1:
            dist.broadcast(ei, 0)
        feats = synth.make_feats_device(ni, dev)
    else:
        ei, feats = synth.make_graph(nu, ni, n_inter, k)
    assert e == int(ei.shape[1])
    note(f"graph generated ({e} edges)")
    # per-head streaming: needed when the per-layer [N, heads*C] tensors kept for the backward do not fit
    n_pad_est = nu + ni + world
    saved_bytes = L * n_pad_est * H * C * (2 if bf16 else 4) if H > 1 else 0
    stream_opt = getattr(args, "stream_heads", "auto")
    stream = (H > 1 and bf16 and cfg["kind"] == "pyg" and not export and
              (stream_opt == "1" or (stream_opt == "auto" and saved_bytes > 0.35 * torch.cuda.get_device_properties(dev).total_memory)))
    tr = ShardedGAT(cfg["kind"], nu, ni, feats, ei, hidden=C, layers=L, heads=H, attn_dropout=0.1, device=dev,
                    feature_dtype=torch.bfloat16 if bf16 else torch.float32, n_triples_max=B.S_TRIPLES, stream_heads=stream)
    if on_device:
        del ei, feats
        ei = feats = None
        torch.cuda.empty_cache()
    note(f"trainer built (rows per rank {tr.n_loc}, heads streamed: {tr.stream})")
    u, i, j = synth.make_triples(nu, ni, B.S_TRIPLES)
    hu, hi, hj = (t.pin_memory() for t in (u, i, j))
    du, di, dj = (t.to(dev) for t in (u, i, j))
    if export:
        step = lambda *_: tr.export_item_embeddings()
    else:
        step = lambda a, b, c: tr.train_step(a, b, c, cfg["loss"])
    sampler = B.ClockSampler(dev.index)
    if rank == 0:
        sampler.start()                   # streams from here on; only the samples inside the timed region are kept
    for w_ in range(args.warmup):
        step(du, di, dj)
        if w_ == 0:
            note("first step done")
    launches0 = _lib.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.begin()
    ev[0].record()
    for _ in range(args.steps):
        step(du, di, dj)
    ev[1].record()
    torch.cuda.synchronize()
    sampler.end()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([ev[0].elapsed_time(ev[1]) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    note(f"timed region done ({float(ms):.2f} ms per step)")
    launches = _lib.launch_count() - launches0
    t_e2e = []
    hout = torch.empty((ni, cfg["hidden"]), dtype=torch.float32).pin_memory() if export else None
    for _ in range(max(min(args.steps, 10) if export else args.steps, 3)):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if export:
            hout.copy_(step(), non_blocking=True)
            torch.cuda.synchronize()
            lv = float(hout[0, 0])
        else:
            uu, ii, jj = (t.to(dev, non_blocking=True) for t in (hu, hi, hj))
            lv = step(uu, ii, jj).item()
        t_e2e.append((time.perf_counter() - t0) * 1e3)
    e2e = torch.tensor([statistics.median(t_e2e)], device=dev)
    if world > 1:
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    # ---- communication: measured on a few extra steps with events around every exchange kernel (wait included)
    if world > 1:
        dist.barrier()            # rank 0 was busy with the clock sampler: start the profiled steps together
    torch.cuda.synchronize()
    tr.fab.stats = {}
    _lib.timing = {}
    n_prof = 5 if e <= 100_000_000 else 2
    for _ in range(n_prof):
        step(du, di, dj)
    torch.cuda.synchronize()
    stats, tr.fab.stats = tr.fab.stats, None
    timing, _lib.timing = _lib.timing, None
    comm_ms = {k_: sum(a.elapsed_time(b) for a, b in v) / n_prof for k_, v in stats.items()}
    kern_ms = {k_: sum(a.elapsed_time(b) for a, b in v) / n_prof for k_, v in timing.items()}
    cm = torch.tensor([sum(comm_ms.values()), sum(kern_ms.values())], device=dev)
    if world > 1:
        dist.all_reduce(cm, op=dist.ReduceOp.MAX)
    comm_bytes = int(tr.comm_bytes_per_step)
    # roofline of this rank's dominant edge kernel: gather-model bytes of ITS rows and edges / its average launch time
    pk, pk_kind = B.peaks()
    rb = 2 if bf16 else 4
    sfx = "bf16" if bf16 else "f32"
    h_alg = 1 if tr.stream else H            # streamed: one head per launch
    alg_f = B.algorithmic_bytes(tr.n_loc, tr.g_fwd.n_edges, h_alg, C, dropout=not export, row_bytes=rb)[f"b200gat_edge_fwd_{sfx}"]
    alg_b = B.algorithmic_bytes(tr.n_loc, tr.g_bwd.n_edges, h_alg, C, dropout=not export, row_bytes=rb)[f"b200gat_edge_bwd_{sfx}"]
    edge_alg = {f"b200gat_edge_fwd_{sfx}": alg_f, f"b200gat_edge_bwd_{sfx}": alg_b, f"b200gat_edge_fwd_stream_{sfx}": alg_f,
                f"b200gat_edge_bwd_phase1_{sfx}": alg_b}
    roof = None
    cand = {k_: v for k_, v in timing.items() if k_ in edge_alg and v}
    if cand:
        dom = max(cand, key=lambda k_: sum(a.elapsed_time(b) for a, b in cand[k_]))
        avg_ms = sum(a.elapsed_time(b) for a, b in cand[dom]) / len(cand[dom])
        ach = edge_alg[dom] / (avg_ms * 1e-3) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": round(ach / pk["hbm_gbs"], 4), "traffic": None, "peak_source": pk_kind,
                "achieved_is": "rank 0: gather-model (algorithmic) bytes of its rows and edges / its average launch time",
                "algorithmic_bytes_per_launch": int(edge_alg[dom]), "avg_launch_ms": round(avg_ms, 4),
                "frac_of_nominal_8TBs": round(ach / 8000.0, 4)}
    parity = None
    if not export and (nu + ni) <= 2_000_000 and not on_device:
        parity = _parity_vs_single(tr, cfg, feats, ei, (du, di, dj), rank, world, dev)
    ms_step, e2e_ms = float(ms), float(e2e)
    if rank == 0:
        conf = B.config_dict(cfg, nu, ni, n_inter, k, world)
        how = ("pushes (every rank stores its block into the peers' buffers with all SMs, then signals)" if tr.fab.mode == "push"
               else "pulls (flag wait + all-SM NVLink reads of the peers' blocks)")
        conf["exchange"] = (f"device-side {how} over peer-mapped memory, no library collective in the step; "
                            + ("layer inputs x exchanged, every rank projects all rows" if tr.x_exchange else "projected rows [h|s] exchanged"))
        conf["rows_per_rank"] = tr.n_loc
        conf["heads_streamed"] = bool(tr.stream)
        conf["graph_generator"] = "torch on the device (synth.make_graph_device)" if on_device else "numpy on the host (synth.make_graph)"
        conf["hbm_peak_allocated_gb_rank0"] = round(torch.cuda.max_memory_allocated(dev) / 2**30, 1)
        print(json.dumps({
            "metric": B.metric_name(cfg), "value": e * L / (ms_step * 1e-3), "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if bf16 else "f32", "data": "synthetic", "config": conf,
            "e2e": {"value": e * L / (e2e_ms * 1e-3), "unit": B.UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 0 if export else int(3 * B.S_TRIPLES * 8),
                    "d2h_bytes_per_step": int(ni * cfg["hidden"] * 4) if export else 4},
            "gpu_launches": int(launches) * world, "clocks": clocks, "roofline": roof, "cpu_baseline": None,
            "epoch_time_ms": ms_step, "loss": lv,
            "comm": {"bytes_pulled_per_rank_per_step": comm_bytes, "ms_per_step_max_rank": round(float(cm[0]), 4),
                     "by_kind_ms_rank0": {k_: round(v, 4) for k_, v in comm_ms.items()},
                     "pull_gbs_rank0": round(comm_bytes / max(comm_ms.get("allgather", 0.0), 1e-9) / 1e6, 1) if world > 1 else None,
                     "compute_kernels_ms_per_step_max_rank": round(float(cm[1]), 4), "overlap_frac": 0.0,
                     "note": "exchange kernels sit in-stream between their producer and their consumer (nothing to overlap them "
                             "with); their time includes waiting for the slowest peer"},
            "breakdown_ms_per_step_rank0": {k_: round(v, 4) for k_, v in sorted(kern_ms.items())},
            "parity_vs_single": parity}))
    if world > 1:
        dist.barrier()
    tr.close()
    if world > 1:
        dist.destroy_process_group()
