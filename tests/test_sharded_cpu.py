"""CPU, world_size 2 over gloo: the host-side logic of the row-sharded path (plan, uneven all-gather, and the
sharding identity itself checked with the oracle standing in for the kernels)."""
import os
import socket

import numpy as np
import pytest
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
        plan = sharded.make_plan(ei, n, rank, world)
        # every rank derives the same bounds
        b = torch.tensor(plan.bounds)
        others = [torch.zeros_like(b) for _ in range(world)]
        dist.all_gather(others, b)
        assert all(torch.equal(o, b) for o in others)
        assert plan.bounds[0] == 0 and plan.bounds[-1] == n and all(x < y for x, y in zip(plan.bounds, plan.bounds[1:]))
        # edge selections: each edge belongs to exactly one rank per direction
        cnt_f = torch.zeros(ei.shape[1], dtype=torch.long); cnt_f[plan.fwd_sel] = 1
        cnt_b = torch.zeros(ei.shape[1], dtype=torch.long); cnt_b[plan.bwd_sel] = 1
        dist.all_reduce(cnt_f); dist.all_reduce(cnt_b)
        assert bool((cnt_f == 1).all()) and bool((cnt_b == 1).all())
        # balance: in+out edge weight of the blocks within 10%
        w = torch.bincount(ei[0], minlength=n) + torch.bincount(ei[1], minlength=n)
        loads = [int(w[plan.bounds[r]:plan.bounds[r + 1]].sum()) for r in range(world)]
        assert max(loads) <= 1.1 * (sum(loads) / world) + w.max().item()
        # uneven all-gather of row blocks
        torch.manual_seed(0)
        full = torch.randn(n, 6)
        got = sharded.all_gather_rows(full[plan.lo:plan.hi].clone(), plan.bounds)
        assert torch.equal(got, full)
        # the sharding identity: a destination block only needs its own in-edges (+ all source rows)
        torch.manual_seed(1)
        c = 16
        x = torch.randn(n, c, dtype=torch.float64)
        W = torch.randn(c, c, dtype=torch.float64) * 0.3
        a_s, a_d = torch.randn(c, dtype=torch.float64), torch.randn(c, dtype=torch.float64)
        y_full = O.simple_gat_layer(x, ei, W, a_s, a_d)
        y_loc = O.simple_gat_layer(x, ei[:, plan.fwd_sel], W, a_s, a_d)[plan.lo:plan.hi]
        np.testing.assert_allclose(y_loc.numpy(), y_full[plan.lo:plan.hi].numpy(), rtol=1e-12, atol=1e-14)
        y_g = sharded.all_gather_rows(y_loc.contiguous(), plan.bounds)
        np.testing.assert_allclose(y_g.numpy(), y_full.numpy(), rtol=1e-12, atol=1e-14)
        # backward identity: a source block's dx only needs its own out-edges once dout / per-destination scalars are shared
        xs = x.clone().requires_grad_(True)
        gy = torch.randn(n, c, dtype=torch.float64)
        (O.simple_gat_layer(xs, ei, W, a_s, a_d) * gy).sum().backward()
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


def test_partition_bounds_edge_cases():
    from b200gat import sharded
    ei = torch.tensor([[0, 1, 2, 3], [3, 2, 1, 0]])
    assert sharded.partition_bounds(ei, 4, 1) == [0, 4]
    b = sharded.partition_bounds(ei, 4, 2)
    assert b == [0, 2, 4]
    hub = torch.stack([torch.zeros(1000, dtype=torch.long), torch.randint(1, 50, (1000,))])
    b = sharded.partition_bounds(hub, 50, 4)
    assert b[0] == 0 and b[-1] == 50 and all(x <= y for x, y in zip(b, b[1:]))
