"""CPU, world_size 2 over gloo: the host-side logic of the row-sharded path (plan, uneven all-gather, and the
sharding identity itself checked with the oracle standing in for the kernels)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from b200gat import sharded, synth
        from oracle import gat_oracle as O
        nu, ni, n_inter, k = synth.CONFIGS["tiny"]
        n = nu + ni
        ei, feats = synth.make_graph(nu, ni, n_inter, k)
        plan = sharded.make_plan(ei, nu, ni, rank, world)
        # the layout is a bijection onto the gathered rows and agrees across ranks
        pm = plan.perm_map
        assert pm.shape == (n,) and len(torch.unique(pm)) == n and int(pm.max()) < world * plan.n_max
        others = [torch.zeros_like(pm) for _ in range(world)]
        dist.all_gather(others, pm)
        assert all(torch.equal(o, pm) for o in others)
        assert torch.equal(pm[plan.local_nodes], plan.lo + torch.arange(plan.n_loc))
        assert plan.n_loc == plan.cu + plan.ci and plan.n_max - plan.n_loc <= 2
        # edge selections: each edge belongs to exactly one rank per direction
        cnt_f = torch.zeros(ei.shape[1], dtype=torch.long); cnt_f[plan.fwd_sel] = 1
        cnt_b = torch.zeros(ei.shape[1], dtype=torch.long); cnt_b[plan.bwd_sel] = 1
        dist.all_reduce(cnt_f); dist.all_reduce(cnt_b)
        assert bool((cnt_f == 1).all()) and bool((cnt_b == 1).all())
        # balance: in- and out-edge counts of the ranks within 10 %
        for sel in (plan.fwd_sel, plan.bwd_sel):
            c = torch.tensor([float(sel.numel())])
            cs = [torch.zeros_like(c) for _ in range(world)]
            dist.all_gather(cs, c)
            loads = [float(x) for x in cs]
            assert max(loads) <= 1.1 * sum(loads) / world, loads
        # padded all-gather of row blocks, then back to node order through perm_map
        torch.manual_seed(0)
        full = torch.randn(n, 6)
        loc = torch.zeros(plan.n_max, 6)
        loc[:plan.n_loc] = full[plan.local_nodes]
        got = sharded.all_gather_rows(loc, world)
        assert got.shape[0] == world * plan.n_max and torch.equal(got[pm], full)
        # the sharding identity: a destination block only needs its own in-edges (+ all source rows), in the
        # permuted row space exactly as the kernels see it
        torch.manual_seed(1)
        c = 16
        x = torch.randn(n, c, dtype=torch.float64)
        W = torch.randn(c, c, dtype=torch.float64) * 0.3
        a_s, a_d = torch.randn(c, dtype=torch.float64), torch.randn(c, dtype=torch.float64)
        y_full = O.simple_gat_layer(x, ei, W, a_s, a_d)
        xp = torch.zeros(world * plan.n_max, c, dtype=torch.float64)
        xp[pm] = x
        y_loc = O.simple_gat_layer(xp, pm[ei][:, plan.fwd_sel], W, a_s, a_d)[plan.lo:plan.lo + plan.n_loc]
        np.testing.assert_allclose(y_loc.numpy(), y_full[plan.local_nodes].numpy(), rtol=1e-12, atol=1e-14)
        pad = torch.zeros(plan.n_max, c, dtype=torch.float64)
        pad[:plan.n_loc] = y_loc
        y_g = sharded.all_gather_rows(pad, world)[pm]
        np.testing.assert_allclose(y_g.numpy(), y_full.numpy(), rtol=1e-12, atol=1e-14)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sharded_plan_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_plan_edge_cases():
    from b200gat import sharded
    ei = torch.tensor([[0, 1, 2, 3], [3, 2, 1, 0]])
    p = sharded.make_plan(ei, 2, 2, 0, 1)
    assert p.n_max == 4 and p.perm_map.tolist() == [0, 1, 2, 3] and p.fwd_sel.tolist() == [0, 1, 2, 3]
    p0, p1 = sharded.make_plan(ei, 2, 2, 0, 2), sharded.make_plan(ei, 2, 2, 1, 2)
    # blocks are padded to a multiple of 4 rows (16-byte aligned per-row blocks of any width): rank 1's block starts at row 4
    assert p0.n_max == 4 and p0.n_loc == 2
    assert p0.perm_map.tolist() == [0, 4, 1, 5] and p0.local_nodes.tolist() == [0, 2] and p1.local_nodes.tolist() == [1, 3]
    # more ranks than users: some ranks own items only
    p3 = sharded.make_plan(torch.zeros((2, 0), dtype=torch.long), 1, 7, 3, 4)
    assert p3.cu == 0 and p3.ci == 1 and p3.n_max == 4
