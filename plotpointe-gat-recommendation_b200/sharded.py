"""Row-sharded full-graph GAT training over P GPUs of one node (one process per GPU, torch.distributed/NCCL).

Nodes are dealt round-robin to the P ranks (user u -> rank u % P, item i -> rank i % P), which balances rows and, for
the reference's graphs, edges.  Rank p owns its nodes' rows: their input features, their destination rows in the
forward (CSR slice) and their source rows in the backward (CSC slice).  Per layer the exchange steps are:

  forward : all-gather of the projected rows [h | s]   (N x (H*C + 2H) floats)      -> fused edge forward on local rows
  backward: all-gather of dout rows and of the per-destination scalars (N x (C + 4H)) -> fused edge backward on local
            source rows (no atomics, no reduce-scatter); all-reduce of the ds_dst partial sums (N x H) and of the
            parameter gradients (< 100 KB)
  loss    : all-gather of the last layer's rows; every rank evaluates the (tiny) triple set and keeps the gradient
            rows of its own block.

The row exchanges pull the peers' blocks out of peer memory with the copy engines (``PeerExchange``, b200gat_peer_* in
include/b200gat.h) when that is the faster path -- 2 ranks by default, B200GAT_PEER=1 forces it, =0 or a box without CUDA IPC
uses NCCL all-gathers, which is also the default beyond 2 ranks (measured, see ShardedGAT.__init__); the small
reductions go through ``torch.distributed`` (plumbing); all arithmetic is the same C-ABI kernels as the single-GPU path.  The plan (row layout, per-rank edge selections) is plain torch and also runs on CPU tensors, which
is how the gloo tests exercise it.
"""
from __future__ import annotations

import json
import os
import statistics
import sys
import time
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------------ the plan
@dataclass
class ShardPlan:
    """Round-robin block layout.  Users u and items i go to rank u % P / i % P, so every rank owns (within one row) the
    same number of users and of items and, for a random graph, the same number of edges.  Inside a block users come
    first, then items (the order node_features produces).  Blocks are padded to ``n_max`` rows so that one
    ``all_gather_into_tensor`` moves a layer; ``perm_map[v]`` is the row of node v in the gathered [P*n_max, ...]
    tensors, and every index array the kernels see (col, row, schedules) is expressed in that row space."""
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
    n_max = max(a + b for a, b in zip(cu_all, ci_all))
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
    """Gather the ranks' padded row blocks ([n_max, ...] each) into one [world*n_max, ...] tensor."""
    if world == 1:
        return local
    local = local.contiguous()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)
    return out


# ------------------------------------------------------------------------------------------------------ peer exchange
class _RawCuda:
    """A raw device allocation seen through __cuda_array_interface__ (so torch can wrap it without copying)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class PeerExchange:
    """All-gather of the ranks' row blocks by pulling them out of peer memory (b200gat_peer_*, include/b200gat.h).

    ``n_buffers`` persistent buffers of ``buffer_bytes`` per rank, one per exchange point of a step; every rank maps all its
    peers' buffers once (CUDA IPC).  ``gather(b, parts)`` = stream-ordered barrier (one-element all-reduce: when it completes
    on this rank's stream, every rank's producer kernels ordered before its own barrier have finished), then one copy-engine
    pull per peer and part on a side stream per peer, then the current stream waits for the pulls.  A buffer is written again
    one step later; the barriers of the other exchange points in between order that write after every peer's pulls."""

    def __init__(self, lib, world: int, rank: int, dev: torch.device, buffer_bytes: int, n_buffers: int):
        import ctypes
        self.lib, self.world, self.rank, self.dev = lib, world, rank, dev
        self.nbytes = (buffer_bytes + 255) // 256 * 256
        self.local_ptr, self.local, handles = [], [], []
        for _ in range(n_buffers):
            p = ctypes.c_void_p()
            lib._check(lib._lib.b200gat_peer_alloc(self.nbytes, ctypes.byref(p)), "peer_alloc")
            h = ctypes.create_string_buffer(64)
            lib._check(lib._lib.b200gat_peer_export(p, h, 64), "peer_export")
            self.local_ptr.append(p.value)
            self.local.append(torch.as_tensor(_RawCuda(p.value, self.nbytes), device=dev))
            handles.append(h.raw)
        gathered = [None] * world
        dist.all_gather_object(gathered, handles)
        self.peer_ptr = []                                   # [rank][buffer] -> device address valid on THIS rank
        for r in range(world):
            if r == rank:
                self.peer_ptr.append(list(self.local_ptr))
                continue
            ptrs = []
            for raw in gathered[r]:
                q = ctypes.c_void_p()
                lib._check(lib._lib.b200gat_peer_open(ctypes.create_string_buffer(raw, 64), ctypes.byref(q)), "peer_open")
                ptrs.append(q.value)
            self.peer_ptr.append(ptrs)
        self.flag = torch.zeros(1, dtype=torch.float32, device=dev)
        self.streams = [torch.cuda.Stream(dev) for _ in range(world - 1)]
        self.order = [(rank + k) % world for k in range(1, world)]       # every rank starts with a different peer

    def view(self, b: int, offset: int, shape, dtype) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= d
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        assert offset % 16 == 0 and offset + nbytes <= self.nbytes
        return self.local[b][offset:offset + nbytes].view(dtype).view(shape)

    def gather(self, b: int, parts) -> list:
        """parts: [(offset, shape_of_one_block, dtype)] -> gathered tensors [(world * rows, ...)]."""
        import ctypes
        lib = self.lib
        dist.all_reduce(self.flag)                           # barrier on the current stream (see class docstring)
        outs, jobs = [], []
        for offset, shape, dtype in parts:
            out = torch.empty((self.world * shape[0],) + tuple(shape[1:]), dtype=dtype, device=self.dev)
            blk = out[:shape[0]].numel() * out.element_size()
            outs.append(out)
            jobs.append((offset, blk, out.data_ptr()))
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        for st, r in zip(self.streams, self.order):
            st.wait_event(ev)
            for offset, blk, dst in jobs:
                lib._check(lib._lib.b200gat_peer_pull(ctypes.c_void_p(dst + r * blk), ctypes.c_void_p(self.peer_ptr[r][b] + offset),
                                                     blk, ctypes.c_void_p(st.cuda_stream)), "peer_pull")
        for offset, blk, dst in jobs:                         # own block: a local copy on the current stream
            lib._check(lib._lib.b200gat_peer_pull(ctypes.c_void_p(dst + self.rank * blk), ctypes.c_void_p(self.local_ptr[b] + offset),
                                                 blk, ctypes.c_void_p(cur.cuda_stream)), "peer_pull")
        for st in self.streams:
            cur.wait_stream(st)
        return outs

    def close(self) -> None:
        for r in range(self.world):
            if r != self.rank:
                for q in self.peer_ptr[r]:
                    self.lib._lib.b200gat_peer_close(q)
        self.peer_ptr = []


# ------------------------------------------------------------------------------------------------------ trainer
class ShardedGAT:
    """Manual forward/backward of the 2-model family (custom / PyG dialect) over a row-sharded graph.

    Parameters are created with the same initialisers and in the same order as the single-GPU modules under the
    same ``torch.manual_seed``, so a sharded run starts from the identical state; ``user_emb`` rows are owned by the
    rank that owns the node, everything else is replicated and its gradient all-reduced."""

    def __init__(self, kind: str, n_users: int, n_items: int, item_feats: torch.Tensor, edge_index: torch.Tensor,
                 hidden: int = 128, layers: int = 2, heads: int = 1, attn_dropout: float = 0.1, seed: int = 42,
                 lr: float = 1e-3, weight_decay: float = 1e-4, device: Optional[torch.device] = None,
                 feature_dtype=torch.float32):
        from . import _lib
        if feature_dtype not in (torch.float32, torch.bfloat16):
            raise NotImplementedError("feature_dtype must be float32 or bfloat16")
        self.bf16 = feature_dtype == torch.bfloat16   # bf16 projection: h and the gathered dout travel (and are stored) as bf16
        from .graph import build_graph
        from .modules import CustomGAT, PyGGAT
        self._lib = _lib
        self.kind = kind
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.nu, self.ni, self.n = n_users, n_items, n_users + n_items
        self.hidden, self.heads, self.n_layers = hidden, (1 if kind == "custom" else heads), layers
        self.p_drop = attn_dropout
        self.training = True
        self.policy = _lib.POLICY_CUSTOM if kind == "custom" else _lib.POLICY_PYG

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
        # node id of every gathered row (-1 for the padding rows): the loss gradient is evaluated for ALL rows on every
        # rank (0.2 ms of redundant work) so that the last layer's dout needs no 354 MB all-gather
        row_nodes = torch.full((self.n_pad,), -1, dtype=torch.int32, device=self.dev)
        row_nodes[plan.perm_map] = torch.arange(self.n, dtype=torch.int32, device=self.dev)
        self.row_nodes = row_nodes
        del ei, ei_p

        torch.manual_seed(seed)
        full = (CustomGAT(n_users, n_items, item_feats.shape[1], hidden, layers) if kind == "custom"
                else PyGGAT(n_users, n_items, item_feats.shape[1], hidden, layers, heads, attn_dropout))
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
        # exchange over peer memory (copy-engine pulls) when the ranks can map each other's buffers; NCCL all-gathers otherwise
        self.px = None
        H_, C_ = self.heads, hidden
        self._offB = (self.n_max * max(H_ * C_, C_) * 4 + 255) // 256 * 256       # part A: h or dout rows, part B: s or nodestat
        # Measured (config 2, fp32): 2 GPUs: exchange 1.22 ms per step with pulls vs 1.86 ms with NCCL; 8 GPUs: 3.6 ms vs 2.4 ms
        # (seven concurrent copy-engine streams per GPU do not add up to the NVLink rate).  So: pulls for 2 ranks, NCCL beyond,
        # unless B200GAT_PEER=1 / 0 forces one of them.
        mode = os.environ.get("B200GAT_PEER", "auto")
        if self.world > 1 and (mode == "1" or (mode == "auto" and self.world <= 2)):
            ok = torch.ones(1, device=self.dev)
            try:
                self.px = PeerExchange(_lib, self.world, self.rank, self.dev, self._offB + self.n_max * 4 * H_ * 4, 2 * layers + 1)
            except Exception as exc:      # noqa: BLE001  (no IPC in this sandbox, no peer access, ...)
                ok.zero_()
                self.px = None
                if self.rank == 0:
                    print(f"b200gat.sharded: peer exchange unavailable ({exc}); using NCCL all-gathers", file=sys.stderr)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)        # all ranks or none
            if ok.item() == 0:
                self.px = None
        self.replicated = list(self.item_proj.parameters()) + self.W + self.a_src + self.a_dst + [b for b in self.bias if b is not None]
        self.opt = torch.optim.Adam([self.user_emb] + self.replicated, lr=lr, weight_decay=weight_decay, fused=True)
        self.step_no = 0
        self.seed = seed

    # -------------------------------------------------------------------------------------------- helpers
    def _empty(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.dev)

    def _rows(self, *shape, dtype=torch.float32):
        """Buffer for a tensor that is exchanged: padded to n_max rows (kernels fill the first n_loc)."""
        return torch.empty((self.n_max,) + shape, dtype=dtype, device=self.dev)

    def _layer_seed(self, layer: int) -> int:
        return (self.seed * 1_000_003 + self.step_no * 101 + layer) & (2 ** 62 - 1)

    # -------------------------------------------------------------------------------------------- forward
    def forward(self) -> torch.Tensor:
        """Returns the local rows of Z; keeps what the backward needs in ``self.saved``."""
        L, H, C, lib = self._lib, self.heads, self.hidden, self._lib
        st = lib.stream()
        from .functional import node_features
        with torch.enable_grad():
            x0 = node_features(self.user_emb, self.item_proj.weight, self.item_proj.bias, self.feats_loc)
        self.x0 = x0
        x = x0.detach()
        self.saved = []
        p = self.p_drop if self.training else 0.0
        for l in range(self.n_layers):
            f_in = x.shape[1]
            h_dt = torch.bfloat16 if self.bf16 else torch.float32
            if self.px is not None:
                h_loc = self.px.view(l, 0, (self.n_max, H * C), h_dt)
                s_loc = self.px.view(l, self._offB, (self.n_max, 2 * H), torch.float32)
            else:
                h_loc, s_loc = self._rows(H * C, dtype=h_dt), self._rows(2 * H)
            dwb = lib.dense_workspace_bytes(H, C, f_in)
            dws = torch.empty(dwb, dtype=torch.uint8, device=self.dev)
            lib.call("b200gat_project_bf16" if self.bf16 else "b200gat_project_f32", lib.ptr(x), lib.ptr(self.W[l]), lib.ptr(self.a_src[l]), lib.ptr(self.a_dst[l]),
                     self.n_loc, f_in, H, C, lib.ptr(h_loc), lib.ptr(s_loc), lib.ptr(dws), dwb, st)
            if self.px is not None:
                h_full, s_full = self.px.gather(l, [(0, (self.n_max, H * C), h_dt), (self._offB, (self.n_max, 2 * H), torch.float32)])
            else:
                h_full = all_gather_rows(h_loc, self.world)
                s_full = all_gather_rows(s_loc, self.world)
            last = l == self.n_layers - 1
            out = self.px.view(self.n_layers, 0, (self.n_max, C), torch.float32) if (self.px is not None and last) else self._rows(C)
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
        return x

    def loss_and_backward(self, z_loc: torch.Tensor, u, i, j, loss_kind: str = "bpr") -> torch.Tensor:
        lib, H, C = self._lib, self.heads, self.hidden
        st = lib.stream()
        L_ = self.n_layers
        if self.px is not None and z_loc.data_ptr() == self.px.local_ptr[L_]:
            z_full = self.px.gather(L_, [(0, (self.n_max, C), torch.float32)])[0]
        else:
            z_full = all_gather_rows(z_loc, self.world)
        s_tr = int(u.shape[0])
        ws_bytes = lib.loss_workspace_bytes(self.n, s_tr)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.dev)
        loss = self._empty(1)
        kind = lib.LOSS_BPR if loss_kind == "bpr" else lib.LOSS_BCE
        lib.call("b200gat_rank_loss_fwd_f32", lib.ptr(z_full), self.nu, self.ni, C, lib.ptr(u), lib.ptr(i), lib.ptr(j), s_tr,
                 lib.ptr(self.node_map), kind, 1, lib.ptr(loss), lib.ptr(ws), ws_bytes, st)
        one = torch.ones(1, dtype=torch.float32, device=self.dev)
        if self.world > 1:
            dout_all = self._empty(self.n_pad, C)
            dout_all_g = torch.empty((self.n_pad, C), dtype=torch.bfloat16, device=self.dev) if self.bf16 else None
            lib.call("b200gat_rank_loss_bwd_f32", lib.ptr(z_full), self.nu, self.ni, C, lib.ptr(u), lib.ptr(i), lib.ptr(j), s_tr,
                     lib.ptr(self.node_map), kind, lib.ptr(one), lib.ptr(self.row_nodes), 0, self.n_pad, lib.ptr(dout_all),
                     lib.ptr(dout_all_g), lib.ptr(ws), ws_bytes, st)
            dout = dout_all[self.plan.lo:self.plan.lo + self.n_max]
            pre_gathered = (dout_all_g if self.bf16 else dout_all)
        else:
            dout = self._rows(C)
            lib.call("b200gat_rank_loss_bwd_f32", lib.ptr(z_full), self.nu, self.ni, C, lib.ptr(u), lib.ptr(i), lib.ptr(j), s_tr,
                     lib.ptr(self.node_map), kind, lib.ptr(one), lib.ptr(self.node_list), 0, self.n_loc, lib.ptr(dout), None,
                     lib.ptr(ws), ws_bytes, st)
            pre_gathered = None
        del z_full
        grads = {}
        for l in reversed(range(self.n_layers)):
            x, h_full, s_full, rowstat, out_h, p, seed = self.saved[l]
            f_in = x.shape[1]
            bx = L_ + 1 + l                                        # this layer's backward exchange buffer
            nodestat = self.px.view(bx, self._offB, (self.n_max, H, 4), torch.float32) if self.px is not None else self._rows(H, 4)
            dwb = lib.dense_workspace_bytes(H, C, f_in)
            dws = torch.empty(dwb, dtype=torch.uint8, device=self.dev)
            db = torch.empty_like(self.bias[l]) if self.bias[l] is not None else None
            have_full = pre_gathered is not None and l == self.n_layers - 1
            dout_g = None
            if self.bf16 and not have_full:
                dout_g = (self.px.view(bx, 0, (self.n_max, C), torch.bfloat16) if self.px is not None
                          else self._rows(C, dtype=torch.bfloat16))
            lib.call("b200gat_node_prep_f32", lib.ptr(dout), lib.ptr(out_h), lib.ptr(self.bias[l] if H == 1 else None),
                     lib.ptr(s_full), lib.ptr(rowstat), self.n_loc, self.plan.lo, H, C, lib.ptr(nodestat), lib.ptr(db),
                     lib.ptr(dout_g), lib.ptr(dws), dwb, st)
            if self.px is not None:
                parts = [] if have_full else [(0, (self.n_max, C), torch.bfloat16 if self.bf16 else torch.float32)]
                if not have_full and not self.bf16:
                    assert dout.data_ptr() == self.px.local_ptr[bx], "fp32 dout of an inner layer must live in its exchange buffer"
                got = self.px.gather(bx, parts + [(self._offB, (self.n_max, H, 4), torch.float32)])
                dout_full = pre_gathered if have_full else got[0]
                nodestat_full = got[-1]
            else:
                dout_full = pre_gathered if have_full else all_gather_rows(dout_g if self.bf16 else dout, self.world)
                nodestat_full = all_gather_rows(nodestat, self.world)
            dh = self._empty(self.n_loc, H * C)
            de = self._empty(max(self.g_bwd.n_edges, 1), H)
            ds = self._empty(self.n_loc, 2 * H)
            sb = self.sched_bwd
            lib.call("b200gat_edge_bwd_bf16" if self.bf16 else "b200gat_edge_bwd_f32", lib.ptr(h_full), lib.ptr(s_full),
                     lib.ptr(dout_full), lib.ptr(nodestat_full),
                     lib.ptr(sb.sched), sb.n_sched, lib.ptr(sb.table), sb.n_long, lib.ptr(sb.partial(H * C + 4)),
                     lib.ptr(self.g_bwd.row), lib.ptr(self.perm_bwd), self.plan.lo, H, C, self.policy, 0.2, lib.ptr(dh),
                     lib.ptr(de), lib.ptr(ds), 2 * H, p, seed, st)
            ds_dst = self._empty(self.n_pad, H)                              # partial sums over this rank's edges
            lib.call("b200gat_ds_dst_f32", lib.ptr(de), lib.ptr(self.g_bwd.rowptr), lib.ptr(self.g_bwd.csr2csc), self.n_pad,
                     self.g_bwd.n_edges, H,
                     lib.ptr(ds_dst), H, st)
            if self.world > 1:
                dist.all_reduce(ds_dst)
            ds[:, H:] = ds_dst[self.plan.lo:self.plan.lo + self.n_loc]
            del de, dout_full, nodestat_full
            # dx of layer l is the dout of layer l-1: in the fp32 tier it is produced straight into that layer's exchange buffer
            if self.px is not None and l >= 1 and not self.bf16:
                dx = self.px.view(L_ + l, 0, (self.n_max, f_in), torch.float32)
            else:
                dx = self._rows(f_in)
            dW, da_s, da_d = torch.empty_like(self.W[l]), torch.empty_like(self.a_src[l]), torch.empty_like(self.a_dst[l])
            lib.call("b200gat_project_bwd_f32", lib.ptr(x), lib.ptr(self.W[l]), lib.ptr(self.a_src[l]), lib.ptr(self.a_dst[l]),
                     lib.ptr(dh), lib.ptr(ds), self.n_loc, f_in, H, C, lib.ptr(dx), lib.ptr(dW), lib.ptr(da_s), lib.ptr(da_d),
                     lib.ptr(dws), dwb, st)
            grads[self.W[l]], grads[self.a_src[l]], grads[self.a_dst[l]] = dW, da_s, da_d
            if db is not None:
                grads[self.bias[l]] = db
            dout = dx
        # input features: user rows take their gradient rows directly, item_proj through torch autograd
        for p_ in [self.user_emb] + list(self.item_proj.parameters()):
            p_.grad = None
        self.x0.backward(dout[:self.n_loc])
        for p_, g in grads.items():
            p_.grad = g
        if self.world > 1:
            flat = torch.cat([p_.grad.reshape(-1) for p_ in self.replicated])
            dist.all_reduce(flat)
            off = 0
            for p_ in self.replicated:
                k = p_.numel()
                p_.grad = flat[off:off + k].view_as(p_)
                off += k
        self.saved = []
        return loss.view(())

    def train_step(self, u, i, j, loss_kind: str = "bpr") -> torch.Tensor:
        z = self.forward()
        loss = self.loss_and_backward(z, u, i, j, loss_kind)
        self.opt.step()
        self.step_no += 1
        return loss

    @torch.no_grad()
    def export_item_embeddings(self) -> torch.Tensor:
        """Config 4: forward-only, returns Z[n_users:] gathered on every rank (tools/export_item_embeddings.py:140-142)."""
        was, self.training = self.training, False
        z = all_gather_rows(self.forward(), self.world)[self.plan.perm_map]     # back to node order
        self.training = was
        self.saved = []
        return z[self.nu:]


# ------------------------------------------------------------------------------------------------------ bench (N > 1)
def bench_main(args, cfg, rank: int, world: int, dev: torch.device) -> None:
    """bench.py's N>1 leg: strong scaling of the same workload, max-over-ranks device time."""
    from . import _lib, synth
    import bench as B
    nu, ni, n_inter, k = B.graph_dims(cfg, synth)
    ei, feats = synth.make_graph(nu, ni, n_inter, k)
    e = int(ei.shape[1])
    bf16 = cfg["tier"] == "bf16"
    export = cfg["mode"] == "export"
    L = cfg["layers"]
    tr = ShardedGAT(cfg["kind"], nu, ni, feats, ei, hidden=cfg["hidden"], layers=L, heads=cfg["heads"], attn_dropout=0.1, device=dev,
                    feature_dtype=torch.bfloat16 if bf16 else torch.float32)
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
    for _ in range(args.warmup):
        step(du, di, dj)
    launches0 = _lib.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.begin()
    ev[0].record()
    for _ in range(args.steps):
        out = step(du, di, dj)
    ev[1].record()
    torch.cuda.synchronize()
    sampler.end()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([ev[0].elapsed_time(ev[1]) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
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
    ms_step, e2e_ms = float(ms), float(e2e)
    if rank == 0:
        conf = B.config_dict(cfg, nu, ni, n_inter, k, world)
        conf["exchange"] = "copy-engine pulls from peer memory" if tr.px is not None else "NCCL all-gather"
        conf["rows_per_rank"] = tr.n_loc
        print(json.dumps({
            "metric": B.metric_name(cfg), "value": e * L / (ms_step * 1e-3), "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if bf16 else "f32", "data": "synthetic", "config": conf,
            "e2e": {"value": e * L / (e2e_ms * 1e-3), "unit": B.UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 0 if export else int(3 * B.S_TRIPLES * 8),
                    "d2h_bytes_per_step": int(ni * cfg["hidden"] * 4) if export else 4},
            "gpu_launches": int(launches) * world, "clocks": clocks, "epoch_time_ms": ms_step, "loss": lv}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
